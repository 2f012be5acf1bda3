import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a CUDA device: the GPU parity tests are skipped (with the reason), not failed.  On a
    GPU box nothing is skipped -- a missing library there is a failure, as it must be."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (the decoder has no CPU path); run with -m gpu on the B200 box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def corpus_files():
    """The 52 compressed streams of the reference's data/ corpus: (name, compressed bytes, expected bytes or None)."""
    out = []
    for fn in sorted(os.listdir(DATA)):
        if ".compressed" not in fn:
            continue
        base = fn.split(".compressed")[0]
        comp = open(os.path.join(DATA, fn), "rb").read()
        if fn.startswith("frewsxcv"):
            exp = None
        else:
            exp = open(os.path.join(DATA, base), "rb").read()
        out.append((fn, comp, exp))
    assert len(out) == 52
    return out


def stream_vectors():
    """The 33 stream tests of the reference's tests/lib.rs plus its doc-test, as (name, input, expected bytes or
    None, expected error substring or None)."""
    vecs = json.load(open(os.path.join(GOLDEN, "stream_vectors.json")))
    out = []
    for v in vecs:
        inp = bytes.fromhex(v["input_hex"]) if "input_hex" in v else open(os.path.join(DATA, v["input_file"]), "rb").read()
        if "expect_error_substring" in v:
            out.append((v["name"], inp, None, v["expect_error_substring"]))
        else:
            exp = bytes.fromhex(v["expect_hex"]) if "expect_hex" in v else open(os.path.join(DATA, v["expect_file"]), "rb").read()
            out.append((v["name"], inp, exp, None))
    return out


# Error classes of the 9 invalid corpus files (the matching tests/lib.rs:397-552 vectors pin the
# message substrings).
FREWSXCV_STATUS = {
    "frewsxcv_01.compressed": 24, "frewsxcv_02.compressed": 8, "frewsxcv_03.compressed": 12,
    "frewsxcv_04.compressed": 1, "frewsxcv_05.compressed": 24, "frewsxcv_06.compressed": 23,
    "frewsxcv_07.compressed": 1, "frewsxcv_08.compressed": 24, "frewsxcv_09.compressed": 10,
}
