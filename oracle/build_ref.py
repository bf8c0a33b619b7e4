"""Stage the REAL reference's hot-path sources under ``oracle/_ref/`` (git-ignored).

    python -m oracle.build_ref          (build container: needs /root/reference)

Test infrastructure (see ``oracle/__init__.py``).  The reference is pure Python, so
"building" it means placing the unmodified source of the path where the GPU box can
run it: ``oracle/_ref/`` travels with the repository snapshot, ``/root/reference``
does not.  Nothing is written outside ``oracle/_ref/`` and nothing from there is ever
committed (``.gitignore``).  What is staged:

* ``multigriddet/postprocess/{nms,wbf,multigrid_decode}.py`` -- byte-for-byte copies;
* ``encoder_functions.py`` -- the four pure-NumPy functions of the target encoder
  (``get_anchor_mask``, ``iol_common_center``, ``best_fit_and_layer``,
  ``preprocess_true_boxes``; generators.py:2473-2544, 3393-3473), each cut out as its
  verbatim source segment (``ast.get_source_segment``): ``generators.py`` as a whole
  executes TensorFlow calls at import and cannot be loaded without it.

``bench.py --impl reference`` times these files (``kind: "reference"``);
``oracle/ref_loader.py`` loads them when ``/root/reference`` is absent.
"""
from __future__ import annotations

import ast
import hashlib
import json
import os
import shutil

from . import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
POST_FILES = ("nms.py", "wbf.py", "multigrid_decode.py")


def build(force: bool = False) -> str | None:
    """Returns the staging directory, or None when the reference tree is not present
    (GPU box: the directory staged in the build container is used as it is)."""
    src_root = os.path.join(ref_loader.REFERENCE_ROOT, "multigriddet")
    gen = os.path.join(src_root, "data", "generators.py")
    if not os.path.isfile(gen):
        return OUT if os.path.isfile(os.path.join(OUT, "MANIFEST.json")) else None
    manifest_path = os.path.join(OUT, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return OUT
    post_out = os.path.join(OUT, "multigriddet", "postprocess")
    os.makedirs(post_out, exist_ok=True)
    manifest = {"source": ref_loader.REFERENCE_ROOT, "files": {}}
    for name in POST_FILES:
        src = os.path.join(src_root, "postprocess", name)
        shutil.copyfile(src, os.path.join(post_out, name))
        with open(src, "rb") as fh:
            manifest["files"]["multigriddet/postprocess/" + name] = hashlib.sha256(fh.read()).hexdigest()
    with open(gen, "r") as fh:
        text = fh.read()
    tree = ast.parse(text, filename=gen)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ref_loader._ENCODER_NAMES]
    if len(picked) != len(ref_loader._ENCODER_NAMES):
        raise RuntimeError("reference encoder functions not found in " + gen)
    with open(os.path.join(OUT, "encoder_functions.py"), "w") as fh:
        fh.write("# verbatim FunctionDef source segments of multigriddet/data/generators.py "
                 "(staged by oracle/build_ref.py; not committed)\nimport numpy as np\n\n\n")
        for n in picked:
            fh.write(ast.get_source_segment(text, n) + "\n\n\n")
            manifest["files"][f"generators.py:{n.name}"] = f"lines {n.lineno}-{n.end_lineno}"
    with open(manifest_path, "w") as fh:
        json.dump(manifest, fh, indent=1)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
