"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, travels to the GPU box with
the gpurun snapshot) so that `bench.py --impl reference` can time the reference's own torch CPU forward there.

    python baseline/stage_reference.py            # build container only: needs /root/reference

The reference is not a pip package (no setup.py / pyproject), so "install" = a verbatim copy of the ~10 Python files its
`models.gwcnet_dca_g.GwcNet` imports.  Nothing is edited; the two import problems of the tree
(models/__init__.py imports a missing file, gwcnet_dca_g.py imports matplotlib) are handled by the shims in
baseline/reference_arm.py, outside the copied files.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DCA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["models/gwcnet_dca_g.py", "models/submodule.py", "models/submodule_bn.py", "models/augment/cva.py",
         "models/augment/semantic_level.py", "models/augment/SelfAttention_bn.py"]
TREES = ["models/lib/nn"]


def stage(verbose=True):
    if not os.path.isdir(REF):
        return False
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    for rel in TREES:
        dst = os.path.join(DST, rel)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, rel), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "tests"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"staged {n} reference files under {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
