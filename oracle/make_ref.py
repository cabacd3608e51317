#!/usr/bin/env python
"""Recipe for ``oracle/_ref/``: a verbatim, git-ignored copy of the reference's own Python files
for the hot path, so that the UNMODIFIED reference can run where /root/reference does not exist
(the GPU box: ``oracle/_ref/`` travels with the snapshot, like the built ``.so`` files).

    python oracle/make_ref.py            # run by __graft_entry__.build() when /root/reference exists

Nothing from the reference is committed: ``oracle/_ref/`` is listed in ``.gitignore``. Only the
checker side uses it -- ``bench.py --impl reference`` / ``cpu_baseline`` (``kind: "reference"``) and
the tests that validate the oracle restatement; no product module imports it. The reference is
pure Python (no build system, nothing to compile); the files keep their package layout because
they import each other as ``encoding.range_image`` etc. with ``src/`` on ``sys.path``.
"""
import os
import shutil
import sys

REFERENCE = os.environ.get("NSC_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

# (file under src/, what it is the reference for)
FILES = [
    ("encoding/__init__.py", "package marker"),
    ("encoding/range_image.py", "RangeImageProjector.project, interpolate_range_image (SURVEY 8 a1-a3)"),
    ("encoding/spectral_encoder.py", "SpectralEncoder (SURVEY 8 a4-a9)"),
    ("encoding/quantization.py", "HistogramQuantizer / CompressedDescriptor (SURVEY 8 f4)"),
    ("retrieval/__init__.py", "package marker"),
    ("retrieval/wasserstein.py", "WassersteinRetriever (SURVEY 8 f1)"),
    ("data/__init__.py", "package marker"),
    ("data/pose_utils.py", "compute_overlap and SE(3) helpers (SURVEY 8 f3)"),
    ("keyframe/__init__.py", "package marker"),
    ("keyframe/criteria.py", "KeyframeSelectionCriteria (SURVEY 8 f3)"),
]


def make(verbose: bool = False) -> bool:
    """Copy the files; returns False (and leaves any existing copy alone) when the reference tree
    is not available, e.g. on the GPU box."""
    src_root = os.path.join(REFERENCE, "src")
    if not os.path.isdir(src_root):
        return False
    for rel, _ in FILES:
        src = os.path.join(src_root, rel)
        dst = os.path.join(DEST, "src", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        elif rel.endswith("__init__.py"):
            open(dst, "w").close()
        else:
            raise FileNotFoundError(src)
        if verbose:
            print("copied", rel)
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Verbatim copy of files of the reference (made by oracle/make_ref.py); git-ignored, never edited.\n")
    return True


def ref_src_path():
    """Directory to put on sys.path to import the reference (``encoding.spectral_encoder`` ...):
    the live tree when present, else the travelling copy, else None."""
    live = os.path.join(REFERENCE, "src")
    if os.path.isdir(os.path.join(live, "encoding")):
        return live
    cp = os.path.join(DEST, "src")
    if os.path.isfile(os.path.join(cp, "encoding", "spectral_encoder.py")):
        return cp
    return None


if __name__ == "__main__":
    ok = make(verbose=True)
    print("oracle/_ref ready" if ok else f"{REFERENCE} not found; nothing copied")
    sys.exit(0)
