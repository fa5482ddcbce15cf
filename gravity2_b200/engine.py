"""Host-side engine: numpy in, numpy out, every computation through the C ABI on a B200.

One ``Engine`` = one ``gv2_ctx`` (one CUDA device + stream).  The module-level
``default_engine()`` is created lazily on first use, never at import time, so that a process
that later forks ``multiprocessing.Pool`` workers (the reference does, upstream of this path:
app/utils/parallel_sig_generator.py:51) has not initialised CUDA just by importing us.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, Optional

import numpy as np

from . import _lib
from ._lib import GravityNativeError, SCHEMES, OUT_DISTANCE, OUT_SIMILARITY  # noqa: F401


def _f64(a, name) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError(f"{name} must be 2-D, got shape {a.shape}")
    return np.ascontiguousarray(a, dtype=np.float64)


def _is_sparse(a) -> bool:
    return hasattr(a, "tocoo") and hasattr(a, "nnz")


def _coo_triplets(a):
    """(rows int32, cols int32, vals float64) of a scipy sparse matrix, explicit zeros dropped, duplicates summed"""
    c = a.tocoo(copy=True)
    c.sum_duplicates()
    c.eliminate_zeros()
    return (np.ascontiguousarray(c.row, dtype=np.int32), np.ascontiguousarray(c.col, dtype=np.int32),
            np.ascontiguousarray(c.data, dtype=np.float64))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, device: Optional[int] = None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", os.environ.get("GRAVITY_B200_DEVICE", "0")))
        h = C.c_void_p()
        rc = self.lib.gv2_create(int(device), C.byref(h))
        if rc != _lib.OK:
            msg = self.lib.gv2_last_error(None)
            raise GravityNativeError(rc, msg.decode() if msg else "?")
        self.ctx = h
        self.device = int(device)
        self._ring = (C.c_void_p(), C.c_void_p(), 0)       # cached (device ring, pinned host ring, bytes)

    def ring(self, nbytes: int):
        """Device + page-locked host staging rings, cached across jobs (pinning memory is slow)."""
        dev, host, have = self._ring
        if have >= nbytes:
            return dev, host
        self._free_ring()
        dev, host = C.c_void_p(), C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, nbytes, C.byref(dev)))
        _lib.check(self.ctx, self.lib.gv2_host_alloc(self.ctx, nbytes, C.byref(host)))
        self._ring = (dev, host, nbytes)
        return dev, host

    def _free_ring(self):
        dev, host, have = self._ring
        if have:
            self.lib.gv2_dev_free(self.ctx, dev)
            self.lib.gv2_host_free(self.ctx, host)
        self._ring = (C.c_void_p(), C.c_void_p(), 0)

    def close(self):
        if getattr(self, "ctx", None):
            self._free_ring()
            self.lib.gv2_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.gv2_launch_count(self.ctx))

    def synchronize(self):
        _lib.check(self.ctx, self.lib.gv2_synchronize(self.ctx))

    # ---------------------------------------------------------------- similarity (a1, a3-a7)
    def similarity(self, sig, gom, aux, scheme: str = "PG", p: float = 1.0, threshold: float = 0.0,
                   apply_threshold: bool = False, distance: bool = False) -> np.ndarray:
        if scheme not in SCHEMES:
            raise ValueError(f"unknown SimilarityMeasurementScheme {scheme!r}")
        need_p = "P" in scheme
        need_g = "G" in scheme
        need_aux = "R" in scheme or "L" in scheme
        sig = _f64(sig, "PPHMMSignatureTable") if need_p else None
        gom = _f64(gom, "GOMSignatureTable") if need_g else None
        aux = _f64(aux, "SPRSignatureTable / location dCor matrix") if need_aux else None
        n = (sig if sig is not None else gom if gom is not None else aux).shape[0]
        for a in (sig, gom, aux):
            if a is not None and a.shape[0] != n:
                raise ValueError("tables disagree on the number of genomes")
        if aux is not None and aux.shape != (n, n):
            raise ValueError(f"aux matrix must be ({n}, {n})")
        out = np.empty((n, n), dtype=np.float64)
        rc = self.lib.gv2_similarity_host(
            self.ctx, _ptr(sig), n, 0 if sig is None else sig.shape[1], 0 if sig is None else sig.shape[1],
            _ptr(gom), 0 if gom is None else gom.shape[1], 0 if gom is None else gom.shape[1], _ptr(aux),
            SCHEMES[scheme], float(p), float(threshold), int(bool(apply_threshold)),
            OUT_DISTANCE if distance else OUT_SIMILARITY, _ptr(out))
        _lib.check(self.ctx, rc)
        return out

    def similarity_neighbourhood(self, sig, cls, weight: float, threshold: float, mul=None) -> np.ndarray:
        """Cleaned PPHMM generalised-Jaccard matrix of the PphmmNeighbourhoodWeight != 0 mode
        (similarity_matrix_constructor.py:112-169 with the visit-count dependent row scaling, quirk Q1)."""
        sig = _f64(sig, "PPHMMSignatureTable")
        n, p = sig.shape
        cls = np.ascontiguousarray(cls, dtype=np.uint8)
        if cls.shape != (n, p):
            raise ValueError("neighbourhood class table must match the signature table")
        mul = None if mul is None else _f64(mul, "multiplier matrix")
        if mul is not None and mul.shape != (n, n):
            raise ValueError(f"multiplier matrix must be ({n}, {n})")
        out = np.empty((n, n), dtype=np.float64)
        rc = self.lib.gv2_similarity_nbhd_host(self.ctx, _ptr(sig), n, p, p, _ptr(cls), float(weight), float(threshold),
                                               _ptr(mul), _ptr(out))
        _lib.check(self.ctx, rc)
        return out

    # --------------------------------------------------------------------- GOM table (a11-a13)
    def gom_table(self, loc, gom_rows, gom_offsets, weights=None, all_columns: bool = False) -> np.ndarray:
        loc = _f64(loc, "PPHMMLocationTable")
        n, p = loc.shape
        offs = np.ascontiguousarray(gom_offsets, dtype=np.int64)
        g = len(offs) - 1
        gom_rows = _f64(gom_rows, "GOM rows") if offs[-1] > 0 else np.zeros((0, p))
        if gom_rows.shape[1] != p or gom_rows.shape[0] != offs[-1]:
            raise ValueError("GOM rows do not match the location table / offsets")
        nb = 1
        w = None
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.int32)
            if w.ndim == 1:
                w = w[None, :]
            if w.shape[1] != p:
                raise ValueError("weights must be (B, P)")
            nb = w.shape[0]
        out = np.zeros((nb, n, g), dtype=np.float64)
        rc = self.lib.gv2_gom_table_host(self.ctx, _ptr(loc), n, p, p, _ptr(gom_rows), p, _ptr(offs), g, _ptr(w), nb,
                                         _lib.GOM_ALL_COLUMNS if all_columns else 0, _ptr(out))
        _lib.check(self.ctx, rc)
        return out if weights is not None else out[0]

    def location_dcor_matrix(self, loc) -> np.ndarray:
        """Scheme "L": dcor of two genomes' location vectors over the union of their hits
        (similarity_matrix_constructor.py:179-186) = GOM table with one single-member GOM per genome."""
        loc = _f64(loc, "PPHMMLocationTable")
        return self.gom_table(loc, loc, np.arange(loc.shape[0] + 1))

    # ------------------------------------------------------------------ presence/absence (a14-a16)
    def shared_counts(self, tab):
        tab = _f64(tab, "signature table")
        n, p = tab.shape
        inter = np.empty((n, n), dtype=np.int32)
        nnz = np.empty((n,), dtype=np.int32)
        rc = self.lib.gv2_shared_counts_host(self.ctx, _ptr(tab), n, p, p, _ptr(inter), _ptr(nnz))
        _lib.check(self.ctx, rc)
        return inter, nnz

    # ------------------------------------------------------------------------- bootstrap (a8, a9)
    def bootstrap(self, sig, loc, gom_of_row, g: int, weights, scheme: str = "PG", p: float = 1.0,
                  distance: bool = True) -> np.ndarray:
        """All replicates at once: (B, N, N) float64.  ``gom_of_row[r]`` = GOM index of loc row r
        (-1 = not a GOMDB member); ``weights`` (B, P) int32 = bincount of each resample index list."""
        if scheme not in ("P", "G", "PG"):
            raise NotImplementedError(f"batched bootstrap covers schemes P, G, PG (got {scheme!r})")
        sig = _f64(sig, "PPHMMSignatureTable")
        n, pp = sig.shape
        same = loc is sig
        loc = sig if same else (_f64(loc, "PPHMMLocationTable") if loc is not None else None)
        w = np.ascontiguousarray(weights, dtype=np.int32)
        if w.ndim == 1:
            w = w[None, :]
        nb = w.shape[0]
        gor = None if gom_of_row is None else np.ascontiguousarray(gom_of_row, dtype=np.int32)
        out = np.empty((nb, n, n), dtype=np.float64)
        rc = self.lib.gv2_bootstrap_host(self.ctx, _ptr(sig), _ptr(loc), n, pp, _ptr(gor), 0 if gor is None else len(gor),
                                         int(g), _ptr(w), nb, SCHEMES[scheme], float(p),
                                         OUT_DISTANCE if distance else OUT_SIMILARITY, _ptr(out))
        _lib.check(self.ctx, rc)
        return out

    # ------------------------------------------------------------------ dendrograms
    def linkage(self, dist, method: str = "average") -> np.ndarray:
        """scipy.cluster.hierarchy.linkage(dist[np.triu_indices_from(dist, k=1)], method) on the device: float64 (n-1, 4)."""
        d = _f64(dist, "distance matrix")
        n = d.shape[0]
        if d.shape != (n, n):
            raise ValueError("distance matrix must be square")
        Z = np.empty((max(n - 1, 0), 4), dtype=np.float64)
        _lib.check(self.ctx, self.lib.gv2_linkage_host(self.ctx, _ptr(d), n, method.encode(), _ptr(Z)))
        return Z

    def newick(self, Z, leaf_names) -> str:
        """GetNewick(to_tree(Z), "", root.dist, leaf_names), app/utils/get_newick.py:1-13 (host, native)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        n = Z.shape[0] + 1
        names = [str(x).encode() for x in leaf_names]
        if len(names) < n:
            raise IndexError("fewer leaf names than leaves")
        arr = (C.c_char_p * len(names))(*names)
        ln = C.c_size_t(0)
        _lib.check(self.ctx, self.lib.gv2_newick_from_linkage(_ptr(Z), n, arr, None, 0, C.byref(ln)))
        buf = C.create_string_buffer(ln.value + 1)
        _lib.check(self.ctx, self.lib.gv2_newick_from_linkage(_ptr(Z), n, arr, buf, ln.value + 1, C.byref(ln)))
        return buf.value.decode()

    def philox_indices(self, seed: int, n_boot: int, p: int) -> np.ndarray:
        out = np.empty((n_boot, p), dtype=np.int64)
        _lib.check(self.ctx, self.lib.gv2_philox_indices_host(self.ctx, C.c_uint64(seed), n_boot, p, _ptr(out)))
        return out


class Job:
    """Device-resident tables (gv2_job): upload once, then the base matrix and any number of bootstrap
    replicates.  Replicates stream through a device ring and a pinned host ring; each one is handed to
    the consumer as a zero-copy (N, N) float64 view that is valid until the consumer asks for the next."""

    def __init__(self, eng: Engine, sig, loc, gom_of_row, g: int, scheme: str = "PG", p: float = 1.0,
                 distance: bool = True, ring_bytes: int = 8 << 30):
        if scheme not in ("P", "G", "PG"):
            raise NotImplementedError(f"device jobs cover schemes P, G, PG (got {scheme!r})")
        self.eng, self.lib, self.ctx = eng, eng.lib, eng.ctx
        self.g = int(g)
        gor = None if gom_of_row is None else np.ascontiguousarray(gom_of_row, dtype=np.int32)
        self.job = C.c_void_p()
        if _is_sparse(sig):
            # scipy sparse tables (the *_coo entries of RefVirusAnnotator.p, app/src/ref_virus_annotator.py:395-400): only the
            # triplets cross PCIe, the dense images are rebuilt in HBM
            self.n, self.p = sig.shape
            sr, sc, sv = _coo_triplets(sig)
            same = loc is sig or loc is None
            if not same and not _is_sparse(loc):
                raise TypeError("sparse signature table with a dense location table: pass both as scipy sparse or both dense")
            lr, lc, lv = (None, None, None) if same or "G" not in scheme else _coo_triplets(loc)
            _lib.check(self.ctx, self.lib.gv2_job_create_coo(
                self.ctx, self.n, self.p, _ptr(sr), _ptr(sc), _ptr(sv), len(sv), _ptr(lr), _ptr(lc), _ptr(lv), 0 if lv is None else len(lv),
                _ptr(gor), 0 if gor is None else len(gor), int(g), SCHEMES[scheme], float(p),
                OUT_DISTANCE if distance else OUT_SIMILARITY, C.byref(self.job)))
        else:
            sig = _f64(sig, "PPHMMSignatureTable")
            self.n, self.p = sig.shape
            loc_arr = sig if (loc is sig or loc is None) else _f64(loc, "PPHMMLocationTable")
            # the same host pointer for sig and loc (PL1, quirk Q3) is uploaded once by the library
            _lib.check(self.ctx, self.lib.gv2_job_create(
                self.ctx, _ptr(sig), _ptr(loc_arr if "G" in scheme else None), 0, self.n, self.p, _ptr(gor),
                0 if gor is None else len(gor), int(g), SCHEMES[scheme], float(p),
                OUT_DISTANCE if distance else OUT_SIMILARITY, C.byref(self.job)))
        self.ntiles = int(self.lib.gv2_num_tiles(self.n))
        mat = max(1, self.n * self.n * 8)
        self.ring_len = max(2, int(ring_bytes // mat))       # two halves: one computing, one draining

    def close(self):
        if getattr(self, "job", None):
            self.lib.gv2_job_destroy(self.job)
            self.job = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def base(self) -> np.ndarray:
        n = self.n
        d = C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, n * n * 8, C.byref(d)))
        try:
            _lib.check(self.ctx, self.lib.gv2_job_base(self.job, 0, self.ntiles, 0, d))
            out = np.empty((n, n))
            _lib.check(self.ctx, self.lib.gv2_memcpy_d2h(self.ctx, _ptr(out), d, n * n * 8))
        finally:
            self.lib.gv2_dev_free(self.ctx, d)
        return out

    def gom_table(self, weights=None) -> np.ndarray:
        """GOM signature table(s) of the job's location table: (N, G) for the un-resampled table, (B, N, G) for
        column multiplicities ``weights`` (B, P) -- what GOMSignatureTable_Constructor returns on ``loc[:, idx]``."""
        n = self.n
        g = self.g
        w = None
        nb = 1
        if weights is not None:
            w = np.ascontiguousarray(weights, dtype=np.int32)
            if w.ndim == 1:
                w = w[None, :]
            if w.shape[1] != self.p:
                raise ValueError("weights must be (B, P)")
            nb = w.shape[0]
        out = np.empty((nb, n, g), dtype=np.float64)
        d_out, d_w = C.c_void_p(), C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, max(out.nbytes, 8), C.byref(d_out)))
        try:
            if w is not None:
                _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, w.nbytes, C.byref(d_w)))
                _lib.check(self.ctx, self.lib.gv2_memcpy_h2d(self.ctx, d_w, _ptr(w), w.nbytes))
            _lib.check(self.ctx, self.lib.gv2_job_gom_table(self.job, d_w if w is not None else None, nb, d_out))
            if out.nbytes:
                _lib.check(self.ctx, self.lib.gv2_memcpy_d2h(self.ctx, _ptr(out), d_out, out.nbytes))
        finally:
            if d_w:
                self.lib.gv2_dev_free(self.ctx, d_w)
            self.lib.gv2_dev_free(self.ctx, d_out)
        return out if weights is not None else out[0]

    def for_each_replicate(self, weights, fn, condensed: bool = False) -> None:
        """fn(b, matrix) for every replicate in order; matrix is a zero-copy view into the pinned ring.
        condensed=True: fn receives scipy's condensed form instead (1-D, n (n - 1) / 2 entries, squareform order -- what
        dist_mat_to_tree.py:8-9 hands to linkage()): half the bytes cross PCIe."""
        w = np.ascontiguousarray(weights, dtype=np.int32)
        if w.ndim == 1:
            w = w[None, :]
        nb = w.shape[0]
        if nb == 0:
            return
        if w.shape[1] != self.p:
            raise ValueError("weights must be (B, P)")
        n = self.n
        ring_len = max(2, min(self.ring_len, nb))
        dev_ring, host_ring = self.eng.ring(ring_len * n * n * 8)
        d_w = C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, w.nbytes, C.byref(d_w)))
        err = []

        m = n * (n - 1) // 2

        def _cb(b, ptr, _user):
            try:
                if condensed:
                    vec = np.frombuffer((C.c_double * m).from_address(ptr), dtype=np.float64) if m else np.zeros(0)
                    fn(int(b), vec)
                else:
                    buf = (C.c_double * (n * n)).from_address(ptr)
                    fn(int(b), np.frombuffer(buf, dtype=np.float64).reshape(n, n))
                return 0
            except BaseException as e:      # never let an exception cross the C frame
                err.append(e)
                return 1

        cb = _lib.REPLICATE_CB(_cb)
        entry = self.lib.gv2_job_replicates_condensed if condensed else self.lib.gv2_job_replicates_stream
        try:
            _lib.check(self.ctx, self.lib.gv2_memcpy_h2d(self.ctx, d_w, _ptr(w), w.nbytes))
            rc = entry(self.job, d_w, nb, dev_ring, ring_len, host_ring, C.cast(cb, C.c_void_p), None)
            if err:
                raise err[0]
            _lib.check(self.ctx, rc)
        finally:
            self.lib.gv2_dev_free(self.ctx, d_w)

    def base_device(self, fn):
        """fn(DeviceMatrix) on the un-resampled matrix, computed into device memory and never shipped."""
        from .classification import DeviceMatrix
        n = self.n
        d = C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, max(8, n * n * 8), C.byref(d)))
        try:
            _lib.check(self.ctx, self.lib.gv2_job_base(self.job, 0, self.ntiles, 0, d))
            self.eng.synchronize()
            return fn(DeviceMatrix(self.eng, d.value, n))
        finally:
            self.lib.gv2_dev_free(self.ctx, d)

    def for_each_replicate_device(self, weights, fn, ring_bytes: int = 16 << 30) -> None:
        """fn(b, DeviceMatrix) for every replicate in order; the matrix stays in HBM (a slot of the device ring) and is
        valid until fn returns -- the consumer reads what it needs with DeviceMatrix.rowmax / .gather."""
        from .classification import DeviceMatrix
        w = np.ascontiguousarray(weights, dtype=np.int32)
        if w.ndim == 1:
            w = w[None, :]
        nb = w.shape[0]
        if nb == 0:
            return
        if w.shape[1] != self.p:
            raise ValueError("weights must be (B, P)")
        n = self.n
        ring_len = max(1, min(nb, int(ring_bytes // max(1, n * n * 8))))
        dev_ring, d_w = C.c_void_p(), C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, ring_len * max(1, n * n * 8), C.byref(dev_ring)))
        err = []

        def _cb(b, ptr, _user):
            try:
                fn(int(b), DeviceMatrix(self.eng, ptr, n))
                return 0
            except BaseException as e:      # never let an exception cross the C frame
                err.append(e)
                return 1

        cb = _lib.REPLICATE_CB(_cb)
        try:
            _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, w.nbytes, C.byref(d_w)))
            _lib.check(self.ctx, self.lib.gv2_memcpy_h2d(self.ctx, d_w, _ptr(w), w.nbytes))
            rc = self.lib.gv2_job_replicates_device(self.job, d_w, nb, dev_ring, ring_len, C.cast(cb, C.c_void_p), None)
            if err:
                raise err[0]
            _lib.check(self.ctx, rc)
        finally:
            if d_w:
                self.lib.gv2_dev_free(self.ctx, d_w)
            self.lib.gv2_dev_free(self.ctx, dev_ring)

    def for_each_tree(self, weights, leaf_names, fn, method: str = "average", ring_bytes: Optional[int] = None) -> None:
        """fn(b, newick) for every replicate in order: the replicate's distance matrix is clustered on the device
        (UPGMA) and printed as DistMat2Tree prints it; the matrices never leave HBM.  ``ring_bytes`` of device memory
        hold the matrices being clustered concurrently (one CTA each; default: up to 128 matrices or 62 % of the
        free memory)."""
        w = np.ascontiguousarray(weights, dtype=np.int32)
        if w.ndim == 1:
            w = w[None, :]
        nb = w.shape[0]
        if nb == 0:
            return
        if w.shape[1] != self.p:
            raise ValueError("weights must be (B, P)")
        n = self.n
        names = [str(x).encode() for x in leaf_names]
        if len(names) < n:
            raise IndexError("fewer leaf names than genomes")
        arr = (C.c_char_p * len(names))(*names)
        mat = max(1, n * n * 8)
        cap = 128
        if ring_bytes is None:
            free_b, total_b = C.c_size_t(0), C.c_size_t(0)
            _lib.check(self.ctx, self.lib.gv2_mem_info(self.ctx, C.byref(free_b), C.byref(total_b)))
            ring_bytes = int(0.62 * free_b.value)
        # four ring parts, three clusterings in flight beside the similarity kernels; measured at 10 000 leaves (lazy column
        # updates): 48 matrices -> 4.70 s, 64 -> 4.01 s, 84 -> 3.45 s, 104 -> 3.42 s, 126 -> 3.20 s per 1001 trees
        ring_len = max(1, min(nb, cap, int(ring_bytes // mat)))
        dev_ring, d_w = C.c_void_p(), C.c_void_p()
        _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, ring_len * mat, C.byref(dev_ring)))
        err = []

        def _cb(b, ptr, ln, _user):
            try:
                fn(int(b), C.string_at(ptr, ln).decode())
                return 0
            except BaseException as e:      # never let an exception cross the C frame
                err.append(e)
                return 1

        cb = _lib.NEWICK_CB(_cb)
        try:
            _lib.check(self.ctx, self.lib.gv2_dev_alloc(self.ctx, w.nbytes, C.byref(d_w)))
            _lib.check(self.ctx, self.lib.gv2_memcpy_h2d(self.ctx, d_w, _ptr(w), w.nbytes))
            rc = self.lib.gv2_job_replicates_newick(self.job, d_w, nb, dev_ring, ring_len, method.encode(), arr,
                                                    C.cast(cb, C.c_void_p), None)
            if err:
                raise err[0]
            _lib.check(self.ctx, rc)
        finally:
            if d_w:
                self.lib.gv2_dev_free(self.ctx, d_w)
            self.lib.gv2_dev_free(self.ctx, dev_ring)

    def replicates(self, weights, condensed: bool = False) -> Iterator[np.ndarray]:
        """Generator form of for_each_replicate: the native call runs on a worker thread and hands each
        matrix over without copying; the GPU keeps computing ahead while the consumer holds a matrix."""
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
            done.wait()                 # consumer has moved on: the ring slot may be reused

        def run():
            try:
                self.for_each_replicate(weights, hand_over, condensed=condensed)
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


_default: Optional[Engine] = None


def default_engine() -> Engine:
    global _default
    if _default is None:
        _default = Engine()
    return _default
