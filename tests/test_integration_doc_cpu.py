"""INTEGRATION.md names only entry points the library exports (four symbols in its table did not exist)."""
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_hg_symbol_named_in_the_docs_is_exported():
    from hgb200._lib import EXPORTED_SYMBOLS
    exported = set(EXPORTED_SYMBOLS)
    header = open(os.path.join(REPO, "include", "hg_api.h")).read()
    declared = set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", header))
    assert declared == exported, (sorted(declared - exported), sorted(exported - declared))
    not_functions = {"hg_api", "hg_conv_desc", "hg_pack_entry", "hg_api_version"}
    for doc in ("INTEGRATION.md", "DESIGN.md", "README.md"):
        text = open(os.path.join(REPO, doc)).read()
        # names of source files (hg_stem.cu, hg_api.h, ...) are not entry points
        for name in set(re.findall(r"\b(hg_[a-z0-9_]+)\b(?!\.(?:cuh|cu|h|o)\b)", text)):
            if name in not_functions or name in exported:
                continue
            assert False, f"{doc} names {name}, which libhgb200 does not export"
