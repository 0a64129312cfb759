"""The entry points named after SURVEY 8(b)'s kernel families (csrc/abi_composites.cu) against the calls they compose:
bit-identical."""
import pytest
import torch

pytestmark = [pytest.mark.gpu]


def test_pool_conv_igemm_and_regress_upsample_equal_their_parts():
    import dcanet_b200 as d
    E, L = d.engine, d._lib
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    st = torch.cuda.current_stream().cuda_stream
    x = E.Planes.from_ncdhw(torch.randn(1, 32, 12, 16, 24, device=dev), 2)
    bn = torch.nn.BatchNorm3d(32).to(dev).eval()
    bn.running_mean.normal_(0, 0.1)
    bn.running_var.uniform_(0.5, 1.5)
    pc = E.PackedConv(torch.randn(32, 32, 3, 3, 3, device=dev) * 0.05, bn)
    assert pc.pack_tc(2)
    E.Options.use_march = False
    try:
        pooled_ref = E.avgpool(x)
        ref = E.conv(pooled_ref, pc, E.K3S1, E.ACT_RELU)
    finally:
        E.Options.use_march = True
    pooled = E.Planes(1, 6, 8, 12, 32, 2, dev)
    y = E.Planes(1, 6, 8, 12, 32, 2, dev)
    L.call("dca_pool_conv", x.ptr, pooled.ptr, pc.w_tc.data_ptr(), pc.scale.data_ptr(), pc.shift.data_ptr(), y.ptr, 2,
           E.ACT_RELU, 1, 32, 12, 16, 24, st)
    y2 = E.Planes(1, 6, 8, 12, 32, 2, dev)
    L.call("dca_conv3d_igemm", E.K3S1, pooled.ptr, 2, pc.w_tc.data_ptr(), pc.scale.data_ptr(), pc.shift.data_ptr(), 0, 0,
           1, 0, 1, 0, 0, y2.ptr, 2, E.ACT_RELU, 1, 32, 32, 6, 8, 12, 6, 8, 12, st)
    torch.cuda.synchronize()
    assert torch.equal(pooled.t, pooled_ref.t) and torch.equal(y.t, ref.t) and torch.equal(y2.t, ref.t)

    logits = torch.randn(1, 12, 16, 24, device=dev)
    mask = torch.randn(1, 1, 16, 24, 144, device=dev)
    pq_ref = E.softmax_regress(logits)
    up_ref = E.convex_upsample(mask, pq_ref)
    pq = torch.empty_like(pq_ref)
    up = torch.empty_like(up_ref)
    L.call("dca_softmax_regress_upsample", logits.data_ptr(), mask.data_ptr(), pq.data_ptr(), up.data_ptr(), 1, 12, 16,
           24, st)
    torch.cuda.synchronize()
    assert torch.equal(pq, pq_ref) and torch.equal(up, up_ref)
