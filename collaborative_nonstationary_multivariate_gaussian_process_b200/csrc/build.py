"""Builds csrc/libnmgp_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "libnmgp_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
PTXAS_OPT = {"nmgp_quadform_mma.cu": os.environ.get("NMGP_PTXAS_QUADFORM", "-O3")}


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found")
    return p


def sources():
    return sorted(glob.glob(os.path.join(HERE, "*.cu")))


INFO = os.path.join(HERE, "BUILD_INFO.json")


def source_digest():
    """SHA-256 over every source / header, this script and the flags: what the built library corresponds to."""
    import hashlib
    h = hashlib.sha256()
    for d in sources() + sorted(glob.glob(os.path.join(HERE, "*.cuh"))) + [os.path.abspath(__file__)]:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    h.update(repr((ARCH, sorted(PTXAS_OPT.items()))).encode())
    return h.hexdigest()


def needs_build():
    """Content-based (not mtime-based): the library is rebuilt whenever a source, header, flag or this script changed."""
    if not os.path.exists(OUT) or not os.path.exists(INFO):
        return True
    try:
        import json
        return json.load(open(INFO)).get("digest") != source_digest()
    except Exception:
        return True


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = nvcc_path()
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        # per-file ptxas level (tuning knob): -O1 keeps the source's k-step-major DMMA order in the quadratic-form kernels
        # where -O3 regroups it into two dependent accumulator chains, but measured 6% slower overall (profiles/README.md)
        popt = PTXAS_OPT.get(os.path.basename(src), "-O3")
        cmd = [nvcc, "-O3", "-std=c++17", "-lineinfo", *ARCH, "-Xcompiler", "-fPIC,-fvisibility=hidden",
               "-Xptxas", popt] + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or "warning" in out:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [nvcc, "-shared", *ARCH, "-o", OUT, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    import json
    import time
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    json.dump({"digest": source_digest(), "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "nvcc": ver,
               "arch": ARCH, "sources": [os.path.basename(x) for x in sources()]}, open(INFO, "w"), indent=1)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
