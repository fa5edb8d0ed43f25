"""Stage the UNMODIFIED reference (junbinhuang/MOFEM) under baseline/_ref/ so that it travels to the GPU box.

    python baseline/stage_ref.py            # build container only: needs /root/reference (read-only)

The reference is pure Python (NumPy / SciPy) with no installer (no setup.py / pyproject), so "installing" it is a copy of
its tree, exactly as SURVEY.md App. B stages it for its own runs: baseline/_ref/MOFEM is the reference, byte for byte;
baseline/_ref/_mplstub is the three-file matplotlib import stub (matplotlib is not installed in this image and is imported
at module level on the geometry and output paths; the stub is ours, tests/golden/_mplstub).  baseline/_ref/ is git-ignored
and NOT gpurun-ignored.  Nothing under mofem_b200/ imports it: only bench.py's CPU-baseline legs (through
baseline/ref_runner.py, in a subprocess) and tests/test_gpu_femmain.py run it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MOFEM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(force=False):
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the reference can only be staged in the build container")
    tree = os.path.join(DST, "MOFEM")
    if os.path.isdir(tree) and not force:
        return tree
    shutil.rmtree(DST, ignore_errors=True)
    shutil.copytree(REF, tree, ignore=shutil.ignore_patterns(".git", "__pycache__"))
    for root, dirs, files in os.walk(tree):          # the source tree is read-only; the copy must be writable (solution.txt etc.)
        for n in dirs + files:
            p = os.path.join(root, n)
            os.chmod(p, os.stat(p).st_mode | 0o200)
    shutil.copytree(os.path.join(os.path.dirname(HERE), "tests", "golden", "_mplstub"), os.path.join(DST, "_mplstub"))
    return tree


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
