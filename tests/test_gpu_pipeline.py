"""HostPipeline (pinned host -> device staging on a copy stream -> hot path -> pinned host) returns what direct calls return."""
import pytest
import torch

from semstereo_b200.params import make_inputs, make_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ORDER = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")


@pytest.mark.parametrize("precision", ["fp32", "bf16", "split"])
def test_pipeline_matches_direct_calls(precision):
    from semstereo_b200.hotpath import DisparityHotPath
    from semstereo_b200.pipeline import HostPipeline
    m = DisparityHotPath(64, False, True, precision=precision)
    m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
    m = m.to(DEV)
    batches = [{k: v.pin_memory() for k, v in make_inputs(30 + i, 1, 128, 128).items()} for i in range(5)]
    direct = [m(*[b[k].to(DEV) for k in ORDER])["pred_up"].cpu() for b in batches]
    pipe = HostPipeline(m, depth=2)
    got = [o.clone() for o in pipe.run(batches)]
    assert len(got) == len(direct)
    for a, b in zip(got, direct):
        assert torch.equal(a, b)
    assert pipe.h2d_bytes == 5 * sum(v.numel() * 4 for v in batches[0].values())
    assert pipe.d2h_bytes == 5 * direct[0].numel() * 4


@pytest.mark.parametrize("precision", ["fp32", "bf16", "split"])
def test_cuda_graph_replay_equals_eager(precision):
    """The whole forward captured as one CUDA graph (every entry point is enqueue-only) reproduces the eager result bit for bit,
    also on new inputs written into the static buffers."""
    from semstereo_b200.graph import GraphedCall
    from semstereo_b200.hotpath import DisparityHotPath
    m = DisparityHotPath(64, False, True, precision=precision)
    m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
    m = m.to(DEV)
    a = {k: v.to(DEV) for k, v in make_inputs(40, 1, 128, 128).items()}
    b = {k: v.to(DEV) for k, v in make_inputs(41, 1, 128, 128).items()}
    call = lambda st: m(*[st[k] for k in ORDER])["pred_up"]       # noqa: E731
    g = GraphedCall(call, a)
    for inp in (a, b, a):
        want = call(inp).clone()
        got = g(inp)
        torch.cuda.synchronize()
        assert torch.equal(got, want)


def test_pipeline_bf16_host_inputs():
    """Host features shipped as bf16 (half the PCIe bytes) are widened on the device: same result as feeding their fp32 values."""
    from semstereo_b200.hotpath import DisparityHotPath
    from semstereo_b200.pipeline import HostPipeline
    m = DisparityHotPath(64, False, True)
    m.load_state_dict(make_params(seed=1, peaked=20.0), strict=True)
    m = m.to(DEV)
    full = [make_inputs(50 + i, 1, 128, 128) for i in range(3)]
    half = [{k: (v.to(torch.bfloat16) if k.startswith("f") else v).pin_memory() for k, v in b.items()} for b in full]     # features bf16, spx / label fp32
    direct = [m(*[b[k].float().to(DEV) for k in ORDER])["pred_up"].cpu() for b in half]
    pipe = HostPipeline(m, depth=2)
    got = [o.clone() for o in pipe.run(half)]
    for a, b in zip(got, direct):
        assert torch.equal(a, b)
    assert pipe.h2d_bytes == 3 * sum(v.numel() * v.element_size() for v in half[0].values())
