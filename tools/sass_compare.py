"""Per-function SASS of the built objects against another commit's build: the proof that moving kernels between files (or
editing emulation-only branches) left the device code instruction for instruction as it was when the GPU suite last ran; the
host side of every object (launch wrappers, C ABI) is compared as x86 disassembly, immediates moved into argument registers
masked (the error-check macros pass __LINE__).
usage: python tools/sass_compare.py <commit>      (builds <commit> in a temporary git worktree; needs nvcc, no GPU)"""
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur is None or re.match(r"^\s*//", line):
            continue
        line = re.sub(r"/\*[0-9a-f]{4}\*/", "", line)             # instruction addresses
        line = re.sub(r"/\* 0x[0-9a-f]* \*/", "", line)           # encodings (they embed absolute branch targets)
        if line.strip():
            out[cur].append(line.strip())
    return out


def host_code(obj):
    """x86 instructions of the host side (launch wrappers, C ABI), addresses stripped; immediates loaded into argument registers
    are masked: the error-check macros pass __LINE__, which moves whenever a kernel moves to a header."""
    txt = subprocess.run(["objdump", "-d", "--no-show-raw-insn", obj], capture_output=True, text=True).stdout
    out = []
    for line in txt.splitlines():
        if "file format" in line:
            continue
        line = re.sub(r"^\s*[0-9a-f]+:\s*", "", line)
        line = re.sub(r"mov\s+\$0x[0-9a-f]+,%(r8d|r9d|ecx|edx|esi|edi)$", r"mov $LINE,%\1", line)
        out.append(line)
    return out


commit = sys.argv[1]
wt = os.path.join(tempfile.mkdtemp(), "wt")
subprocess.run(["git", "-C", ROOT, "worktree", "add", "-q", wt, commit], check=True)
try:
    subprocess.run(["make", "-C", os.path.join(wt, "gravity2_b200", "csrc"), "-j4"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", os.path.join(ROOT, "gravity2_b200", "csrc")], check=True, stdout=subprocess.DEVNULL)
    total = diff = hdiff = 0
    for o in sorted(glob.glob(os.path.join(wt, "gravity2_b200", "csrc", "*.o"))):
        a, b = functions(o), functions(os.path.join(ROOT, "gravity2_b200", "csrc", os.path.basename(o)))
        total += len(a)
        for k in sorted(set(a) | set(b)):
            if a.get(k) != b.get(k):
                diff += 1
                print(f"{os.path.basename(o)}: {k} differs")
        if host_code(o) != host_code(os.path.join(ROOT, "gravity2_b200", "csrc", os.path.basename(o))):
            hdiff += 1
            print(f"{os.path.basename(o)}: host code differs")
    print(f"{total} device functions compared with {commit}: {diff} differ; host code of {hdiff} objects differs "
          f"(__LINE__ immediates masked)")
finally:
    subprocess.run(["git", "-C", ROOT, "worktree", "remove", "--force", wt])
sys.exit(1 if diff or hdiff else 0)
