"""TEST / BASELINE INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference sources the parity tests and the baseline arms
of bench.py execute into `oracle/_ref/`, so that they travel to the GPU box with the gpurun snapshot
(`/root/reference` exists only in the build container):

    python -m oracle.build_ref            # also run by __graft_entry__.build() when /root/reference is present

`oracle/_ref/` is git-ignored (reference sources never enter this repository's history) but not gpurun-ignored.
Files staged (verbatim copies, no edits -- a SHA-256 manifest is written beside them):
    AVE/model/Swin_AVE.py, AVE/model/CLIP_AVE.py     the reference modules (the oracle's ground truth, the eager arm)
    AVE/traintest_adapt_ave29.py                      the unchanged caller (train / validate loops)
    utilities/{__init__,util,stats,scheduler}.py      what the caller star-imports
Consumers: oracle/shim.py::load_reference (falls back to oracle/_ref when /root/reference is absent),
bench.py --impl reference / --impl eager, tests/test_reference_gpu.py.
"""
import hashlib
import json
import os
import shutil

SRC = os.environ.get("STGCMA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ["AVE/model/Swin_AVE.py", "AVE/model/CLIP_AVE.py", "AVE/traintest_adapt_ave29.py", "utilities/__init__.py",
         "utilities/util.py", "utilities/stats.py", "utilities/scheduler.py"]


def main():
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        print(f"oracle.build_ref: {SRC} not present; nothing staged (an existing oracle/_ref is left as is)")
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    print(f"oracle.build_ref: staged {len(FILES)} unmodified reference files in {DST}")
    return True


if __name__ == "__main__":
    main()
