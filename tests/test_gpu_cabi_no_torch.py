"""The C ABI without torch anywhere: device memory and the stream come from the CUDA runtime (cuda-python), the library is
bound with ctypes, inputs and the check are numpy.  What a non-torch host of the reference's path would do
(INTEGRATION.md section 4); proves that no entry point needs a torch type, allocator or stream."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "cost-volume-aggregation-in-stereo-matching-revisited_b200", "libdca_b200.so")


def _rt():
    from cuda.bindings import runtime as rt
    return rt


def _ok(res):
    err = res[0]
    assert int(err) == 0, err
    return res[1] if len(res) == 2 else res[1:]


def test_gwc_volume_and_regression_through_ctypes_and_the_cuda_runtime():
    rt = _rt()
    lib = ctypes.CDLL(LIB)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.dca_build_gwc_volume_f32.argtypes = [vp, vp, vp] + [ci] * 6 + [vp]
    lib.dca_softmax_regress.argtypes = [vp, vp] + [ci] * 4 + [vp]
    B, C, G, D, H, W = 1, 320, 40, 12, 6, 40
    rng = np.random.default_rng(0)
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    stream = _ok(rt.cudaStreamCreate())
    dL, dR = _ok(rt.cudaMalloc(L.nbytes)), _ok(rt.cudaMalloc(R.nbytes))
    vol = np.empty((B, G, D, H, W), np.float32)
    dV = _ok(rt.cudaMalloc(vol.nbytes))
    pred = np.empty((B, 1, H, W), np.float32)
    dP = _ok(rt.cudaMalloc(pred.nbytes))
    try:
        _ok(rt.cudaMemcpyAsync(dL, L.ctypes.data, L.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, stream))
        _ok(rt.cudaMemcpyAsync(dR, R.ctypes.data, R.nbytes, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, stream))
        assert lib.dca_build_gwc_volume_f32(int(dL), int(dR), int(dV), B, C, G, D, H, W, int(stream)) == 0
        # the first group of the volume doubles as [B, D, H, W] logits for the fused softmax + regression
        assert lib.dca_softmax_regress(int(dV), int(dP), B, D, H, W, int(stream)) == 0
        _ok(rt.cudaMemcpyAsync(vol.ctypes.data, dV, vol.nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, stream))
        _ok(rt.cudaMemcpyAsync(pred.ctypes.data, dP, pred.nbytes, rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, stream))
        _ok(rt.cudaStreamSynchronize(stream))
    finally:
        for p in (dL, dR, dV, dP):
            rt.cudaFree(p)
        rt.cudaStreamDestroy(stream)
    # numpy restatement of build_gwc_volume (submodule.py:148-167) and softmax + disparity_regression (:127-131)
    ref = np.zeros_like(vol)
    cpg = C // G
    for d in range(D):
        prod = (L[:, :, :, d:] * R[:, :, :, :W - d]) if d else (L * R)
        ref[:, :, d, :, d:] = prod.reshape(B, G, cpg, H, W - d).mean(axis=2)
    assert np.abs(vol - ref).max() < 1e-5
    logits = ref[:, 0]
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    want = (p * np.arange(D, dtype=np.float32).reshape(1, D, 1, 1)).sum(axis=1, keepdims=True)
    assert np.abs(pred - want).max() < 1e-4
    # a bad call returns the status code instead of throwing or crashing
    assert lib.dca_build_gwc_volume_f32(0, int(0), int(0), B, C, G, D, H, W, 0) == -1
