"""
Recipe for oracle/_ref -- TEST / BENCHMARK INFRASTRUCTURE ONLY.

Makes the UNMODIFIED reference importable where /root/reference does not exist
(the GPU box): copies the reference's Python package (code/src, code/model_parameters,
code/__init__.py) from where it lies into the git-ignored directory oracle/_ref/.
Nothing is edited; oracle/_ref/ is listed in .gitignore (never committed) and
not in .gpurunignore (it travels with the snapshot).  Run by
__graft_entry__.build() whenever /root/reference is present.

Used by: bench.py --impl reference and bench.py's cpu_baseline leg (through
oracle/ref_runner.py), which time the reference's own Simulation.run on the
box's host cores.  The product (hydromodel_b200/) never touches it.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("HB_REFERENCE_ROOT", "/root/reference/code")
REF_DST = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF_DST, "src"))


def make(force=False):
    """Copy the reference package into oracle/_ref (idempotent).  Returns the
    directory, or None when the reference is not present on this machine."""
    if not os.path.isdir(os.path.join(REF_SRC, "src")):
        return REF_DST if available() else None
    if available() and not force:
        return REF_DST
    if os.path.isdir(REF_DST):
        shutil.rmtree(REF_DST)
    os.makedirs(REF_DST)
    for name in ("src", "model_parameters"):
        shutil.copytree(os.path.join(REF_SRC, name), os.path.join(REF_DST, name),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for name in ("__init__.py", "berkeley_hydro_main.py"):
        if os.path.exists(os.path.join(REF_SRC, name)):
            shutil.copy(os.path.join(REF_SRC, name), os.path.join(REF_DST, name))
    return REF_DST


if __name__ == "__main__":
    print(make(force=True))
