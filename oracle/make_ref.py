"""Recipe for oracle/_ref/: the UNMODIFIED upstream module the reference arm of bench.py times.

    python oracle/make_ref.py        (also run by __graft_entry__.build() when /root/reference is present)

The upstream project is pure Python, so "building" the reference is placing its own file where the GPU box can import
it: /root/reference does not exist there, but oracle/_ref/ travels with the repository snapshot (it is git-ignored --
never part of this repository's history -- and not gpurun-ignored, like the built .so files).  Only the module on the
measured path is staged: modules/svd_linear.py (SVDLinear.from_linear, upstream modules/svd_linear.py:26-103).
Nothing under asvd4llm_b200/ imports oracle/_ref; tests/ and bench.py's CPU legs are the only users."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ASVD_REFERENCE", "/root/reference")
FILES = ["modules/svd_linear.py"]


def make(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/_ref: {REF} not present; keeping whatever is staged already")
        return False
    for rel in FILES:
        dst = os.path.join(HERE, "_ref", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        init = os.path.join(os.path.dirname(dst), "__init__.py")
        if not os.path.exists(init):
            open(init, "w").close()
    open(os.path.join(HERE, "_ref", "__init__.py"), "w").close()
    if verbose:
        print("oracle/_ref staged from", REF)
    return True


def load_upstream_svd_linear():
    """upstream's modules.svd_linear module imported from oracle/_ref, or None when it was never staged."""
    path = os.path.join(HERE, "_ref", "modules", "svd_linear.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_asvd_upstream_svd_linear", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
