"""Every GPU of the node behind ONE caller process: ``MultiEngine`` / ``MultiJob`` mirror
``Engine`` / ``Job`` on top of the library's ``gv2_multi_*`` entry points -- one host thread and one context per B200
inside libgravity_b200, tables uploaded once and broadcast over NVLink, bootstrap replicates dealt to the GPUs by index,
the un-resampled matrix as tile ranges all-gathered with NCCL.  An unmodified pipeline_i / pipeline_ii process (one
Python thread) gets all GPUs through ``bootstrap_newick`` / ``bootstrap_distance_matrices``:

    GRAVITY_B200_DEVICES=all   (or "0,1,2,3")   python -m app.pipeline_i ...
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import GravityNativeError, SCHEMES, OUT_DISTANCE, OUT_SIMILARITY


def _check(lib, m, rc):
    if rc != _lib.OK:
        msg = lib.gv2_multi_last_error(m)
        raise GravityNativeError(rc, msg.decode() if msg else "?")


class MultiEngine:
    def __init__(self, devices: Optional[Sequence[int]] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.gv2_multi_create(None, 0, C.byref(h))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.gv2_multi_create(arr, len(devices), C.byref(h))
        if rc != _lib.OK:
            msg = self.lib.gv2_multi_last_error(None)
            raise GravityNativeError(rc, msg.decode() if msg else "?")
        self.m = h
        self.n_devices = int(self.lib.gv2_multi_device_count(h))
        self.uses_nccl = bool(self.lib.gv2_multi_uses_nccl(h))

    @property
    def launches(self) -> int:
        return sum(int(self.lib.gv2_launch_count(C.c_void_p(self.lib.gv2_multi_context(self.m, i)))) for i in range(self.n_devices))

    def close(self):
        if getattr(self, "m", None):
            self.lib.gv2_multi_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiJob:
    """Same surface as engine.Job (base, for_each_tree, for_each_replicate, replicates, close) on all GPUs of a MultiEngine."""

    def __init__(self, meng: MultiEngine, sig, loc, gom_of_row, g: int, scheme: str = "PG", p: float = 1.0, distance: bool = True):
        if scheme not in ("P", "G", "PG"):
            raise NotImplementedError(f"device jobs cover schemes P, G, PG (got {scheme!r})")
        self.meng, self.lib = meng, meng.lib
        sig = np.ascontiguousarray(np.asarray(sig), dtype=np.float64)
        self.n, self.p = sig.shape
        same = loc is None or loc is sig
        loc_arr = sig if same else np.ascontiguousarray(np.asarray(loc), dtype=np.float64)
        gor = None if gom_of_row is None else np.ascontiguousarray(gom_of_row, dtype=np.int32)
        self._keep = (sig, loc_arr, gor)
        self.job = C.c_void_p()
        vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _check(self.lib, meng.m, self.lib.gv2_multi_job_create(
            meng.m, vp(sig), vp(loc_arr if "G" in scheme else None), self.n, self.p, vp(gor), 0 if gor is None else len(gor), int(g),
            SCHEMES[scheme], float(p), OUT_DISTANCE if distance else OUT_SIMILARITY, C.byref(self.job)))

    def close(self):
        if getattr(self, "job", None):
            self.lib.gv2_multi_job_destroy(self.job)
            self.job = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def base(self) -> np.ndarray:
        out = np.empty((self.n, self.n), dtype=np.float64)
        _check(self.lib, self.meng.m, self.lib.gv2_multi_job_base_host(self.job, out.ctypes.data_as(C.c_void_p)))
        return out

    def _weights(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.int32)
        if w.ndim == 1:
            w = w[None, :]
        if w.shape[1] != self.p:
            raise ValueError("weights must be (B, P)")
        return w

    def for_each_tree(self, weights, leaf_names, fn, method: str = "average", ring_len: int = 0) -> None:
        w = self._weights(weights)
        if w.shape[0] == 0:
            return
        names = [str(x).encode() for x in leaf_names]
        if len(names) < self.n:
            raise IndexError("fewer leaf names than genomes")
        arr = (C.c_char_p * len(names))(*names)
        err = []

        def _cb(b, ptr, ln, _user):
            try:
                fn(int(b), C.string_at(ptr, ln).decode())
                return 0
            except BaseException as e:      # never let an exception cross the C frame
                err.append(e)
                return 1

        cb = _lib.NEWICK_CB(_cb)
        rc = self.lib.gv2_multi_job_replicates_newick(self.job, w.ctypes.data_as(C.c_void_p), w.shape[0], method.encode(), arr,
                                                      int(ring_len), C.cast(cb, C.c_void_p), None)
        if err:
            raise err[0]
        _check(self.lib, self.meng.m, rc)

    def for_each_replicate(self, weights, fn, ring_len: int = 0) -> None:
        w = self._weights(weights)
        if w.shape[0] == 0:
            return
        n = self.n
        err = []

        def _cb(b, ptr, _user):
            try:
                buf = (C.c_double * (n * n)).from_address(ptr)
                fn(int(b), np.frombuffer(buf, dtype=np.float64).reshape(n, n))
                return 0
            except BaseException as e:
                err.append(e)
                return 1

        cb = _lib.REPLICATE_CB(_cb)
        rc = self.lib.gv2_multi_job_replicates_host(self.job, w.ctypes.data_as(C.c_void_p), w.shape[0], int(ring_len),
                                                    C.cast(cb, C.c_void_p), None)
        if err:
            raise err[0]
        _check(self.lib, self.meng.m, rc)

    def replicates(self, weights) -> Iterator[np.ndarray]:
        """Generator form of for_each_replicate (the native call runs on a worker thread, matrices are handed over
        without copying)."""
        import queue
        import threading
        q: "queue.Queue" = queue.Queue(maxsize=1)
        done = threading.Event()
        stop = []

        def hand_over(b, m):
            if stop:
                raise StopIteration
            done.clear()
            q.put((b, m))
            done.wait()

        def run():
            try:
                self.for_each_replicate(weights, hand_over)
                q.put(None)
            except StopIteration:
                q.put(None)
            except BaseException as e:
                q.put(e)

        t = threading.Thread(target=run, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                yield item[1]
                done.set()
        finally:
            stop.append(True)
            done.set()
            t.join()


_default_multi: Optional[MultiEngine] = None


def devices_from_env() -> Optional[list]:
    """GRAVITY_B200_DEVICES = "all" | "0,1,2": the GPUs the batched bootstrap generators use by default (unset: one GPU)."""
    spec = os.environ.get("GRAVITY_B200_DEVICES", "").strip()
    if not spec:
        return None
    if spec.lower() == "all":
        return list(range(int(_lib.load().gv2_device_count())))
    return [int(x) for x in spec.split(",") if x.strip() != ""]


def default_multi_engine() -> Optional[MultiEngine]:
    global _default_multi
    devs = devices_from_env()
    if not devs or len(devs) < 2:
        return None
    if _default_multi is None:
        _default_multi = MultiEngine(devs)
    return _default_multi
