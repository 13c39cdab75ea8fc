"""The drop-in boundary on a machine without a GPU: libhgb200.so loads, exports every symbol include/hg_api.h declares
(and the ctypes binding declares nothing else), reports the header's API version, and its argument validation answers
with an error code + message instead of touching a device.  No compute calls."""
import ctypes as C
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "hg_api.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return sorted(set(re.findall(r"\b(?:int|int64_t|size_t)\s+(hg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree():
    from hgb200._lib import lib, lib_path, EXPORTED_SYMBOLS
    names = declared_symbols()
    assert len(names) >= 35 and "hg_conv_nhwc_bf16" in names and "hg_decode_final_preds_v2" in names
    raw = C.CDLL(lib_path())
    for n in names:
        assert hasattr(raw, n), f"{n} is declared in include/hg_api.h but not exported by {lib_path()}"
    assert sorted(EXPORTED_SYMBOLS) == names, (sorted(set(EXPORTED_SYMBOLS) ^ set(names)))
    version = int(re.search(r"#define\s+HG_API_VERSION\s+(\d+)", open(HEADER).read()).group(1))
    assert lib.hg_api_version() == version


def test_argument_validation_needs_no_device():
    from hgb200._lib import lib, HgError
    assert lib.hg_conv_nhwc_bf16(None, None) != 0
    buf = C.create_string_buffer(256)
    assert lib.hg_last_error(buf, 256) > 0 and b"descriptor" in buf.value
    assert lib.hg_decode_argmax(None, None, None, None, 1, 1, 4, 4, None) != 0
    assert lib.hg_dwconv3x3_nhwc(None, None, None, None, 1, 4, 4, 64, 0, 0, None) != 0
    assert lib.hg_preprocess_frames_u8(None, None, None, None, 1, 4, 4, 4, 4, None) != 0
    with pytest.raises(HgError):
        lib.check(lib.hg_wgrad_bf16(None, None, None, None, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, None), "hg_wgrad_bf16")
    assert lib.hg_colstats_nhwc(None, None, None, 0, 64, 64, 0, None, None) != 0
    assert lib.hg_colreduce_scratch_bytes(0, 64) == 0 and lib.hg_colreduce_scratch_bytes(64, 48) == 0
    # the windowed stem: tiles are runs of 128 output pixels of ONE output row; a width whose rows do not split into whole
    # tiles is refused (it would read past the packed row and store past the output row), before any CUDA call
    dummy = C.c_void_p(16)
    assert lib.hg_stem_conv(dummy, dummy, dummy, dummy, None, 1, 64, 384, None) != 0
    assert lib.hg_last_error(buf, 256) > 0 and b"multiples of 256" in buf.value


def test_product_path_fails_loudly_without_cuda():
    """No CPU fallback anywhere on the product path: CPU tensors are refused (tests/test_*_cpu.py cover the models)."""
    import torch
    from hgb200 import ops, HgError
    with pytest.raises(HgError):
        ops.maxpool2x2(torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16))
    with pytest.raises(HgError):
        ops.decode_argmax(torch.zeros(1, 1, 4, 4))
    with pytest.raises(HgError):
        ops.normalize_u8(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), [0, 0, 0], [1, 1, 1])


def test_weights_key_of_the_model_sees_every_kind_of_update_and_is_cheap():
    """HourglassNet._weights_key decides whether the folded inference engine is still valid; it runs on EVERY forward, so it
    must be cheap (it used to cost more than the batch-1 forward itself) and must change on in-place updates, on
    load_state_dict, on a replaced Parameter and on the fused training step's epoch counter."""
    import time
    import torch
    from src.models import hg
    model = hg(num_stacks=8, num_blocks=1, num_classes=16, mobile=False, skip_mode="sum")
    k0 = model._weights_key("cuda:0")
    assert model._weights_key("cuda:0") == k0 and model._weights_key("cuda:1") != k0
    with torch.no_grad():
        model.score[3].bias.add_(1.0)
    k1 = model._weights_key("cuda:0")
    assert k1 != k0
    model.hg[2].hg[0][0][0].bn1.running_mean.mul_(0.5)                       # a buffer
    k2 = model._weights_key("cuda:0")
    assert k2 != k1
    model.load_state_dict({k: v.clone() for k, v in model.state_dict().items()})
    k3 = model._weights_key("cuda:0")
    assert k3 != k2
    model.fc[0][0].weight = torch.nn.Parameter(model.fc[0][0].weight.detach().clone())     # a replaced Parameter object
    k4 = model._weights_key("cuda:0")
    assert k4 != k3
    model._weights_epoch = 5
    assert model._weights_key("cuda:0") != k4
    t0 = time.perf_counter()
    for _ in range(20):
        model._weights_key("cuda:0")
    per_call = (time.perf_counter() - t0) / 20
    assert per_call < 1e-3, per_call
    print(f"_weights_key: {per_call * 1e6:.0f} us per call")
