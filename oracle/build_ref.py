"""Vendors the UNMODIFIED reference sources of the hot path into oracle/_ref/ (git-ignored; it travels to the GPU
box with the gpurun snapshot like a built .so) so that ``bench.py --impl reference`` and the ``cpu_baseline`` leg time
the real reference there instead of the oracle port.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under the product package imports this.  The files are copied
byte for byte (a manifest with their sha256 is written next to them); the three torch>=2 compatibility patches of
SURVEY.md 8c are applied at import time by ``oracle/ref_runner.py``, never to the files.

    python oracle/build_ref.py            # no-op (returns False) when /root/reference is absent
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = [
    "LICENSE",
    "code/nmgp_dsvi.py", "code/utils.py", "code/pre_nmgp.py",
    "code/SIM_code/Utility/kernels.py", "code/SIM_code/Utility/kronecker_operation.py",
    "code/SIM_code/Utility/distributions.py", "code/SIM_code/Utility/settings.py",
    "code/SIM_code/Utility/utils.py", "code/SIM_code/Utility/logpos.py", "code/SIM_code/Utility/prediction.py",
]


def build(force=False):
    if not os.path.isdir(REF):
        return os.path.isdir(OUT)
    man = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        data = open(src, "rb").read()
        man[rel] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": man}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    print("oracle/_ref", "ready" if build(force=True) else "unavailable (no /root/reference)")
