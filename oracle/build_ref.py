"""Recipe for oracle/_ref: the UNMODIFIED reference, byte-compiled from its sources where they lie under /root/reference.

Test / measurement infrastructure only (like everything under oracle/): the product never imports it.

The reference is pure Python (no native code), so "building" it means `py_compile` of the four files of the hot
path into oracle/_ref/modules/*.pyc (sourceless import layout).  No reference source text enters this repository:
oracle/_ref/ is git-ignored, the .pyc files are build outputs that travel to the GPU box with the working tree, where
`bench.py --impl reference` and the `cpu_baseline` leg import them (`from modules.xfeat import XFeat`) and drive the
reference's own public API on the host CPU.  Weights come from this repo's weights/xfeat_state.npz (the published
state_dict re-saved), passed to the constructor as a dict (modules/xfeat.py:29-35 accepts one).

    python oracle/build_ref.py            # needs /root/reference; prints the output directory
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("XFEAT_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
FILES = ["modules/__init__.py", "modules/xfeat.py", "modules/model.py", "modules/interpolator.py"]


def build_ref(force: bool = False) -> str | None:
    """Compile the reference modules into oracle/_ref. Returns the directory, or None when /root/reference is absent
    (the GPU box: the prebuilt files are used as they are)."""
    if not os.path.isdir(REF_ROOT):
        return OUT if available() else None
    for rel in FILES:
        src = os.path.join(REF_ROOT, rel)
        dst = os.path.join(OUT, rel + "c")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
    return OUT


def available() -> bool:
    return all(os.path.exists(os.path.join(OUT, rel + "c")) for rel in FILES)


def import_reference():
    """-> the reference's XFeat class (modules.xfeat.XFeat), imported from oracle/_ref. CPU use: export CUDA_VISIBLE_DEVICES=''
    before torch is imported (the reference picks CUDA when it sees one, modules/xfeat.py:25)."""
    if not available():
        raise RuntimeError("oracle/_ref is not built (run `python oracle/build_ref.py` where /root/reference exists)")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from modules.xfeat import XFeat  # type: ignore
    return XFeat


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
