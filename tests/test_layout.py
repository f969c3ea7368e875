"""CPU-side checks: the C-ABI library builds, loads and exports every declared symbol; the product never imports the
oracle; state_dict compatibility; sharding logic incl. a world_size-2 gloo run; reference arm of bench.py."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from semstereo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "semstereo_b200.h")).read()
    declared = set(re.findall(r"^SS_API\s+[\w\s\*]+?\b(ss_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                      # builds with nvcc if needed (cross-compiles without a GPU)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ss_version() >= 100
    assert lib.ss_ssr_param_count(6) == 189
    assert _lib.last_error() == ""


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "semstereo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_no_cpu_fallback():
    from semstereo_b200 import ops, submodule
    z = torch.zeros(1, 8, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gwc_volume(z, z, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        submodule.build_concat_volume(z, z, 2)


def test_state_dict_matches_reference_inventory():
    from semstereo_b200.hotpath import DisparityHotPath, pack_ssr
    from semstereo_b200.params import hotpath_param_shapes, make_params
    m = DisparityHotPath(64)
    sd = {k: v for k, v in m.state_dict().items() if not k.endswith("num_batches_tracked")}
    shapes = hotpath_param_shapes()
    assert set(sd) == set(shapes)
    assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in shapes)
    p = make_params(1)
    m.load_state_dict({"module." + k: v for k, v in p.items()}, strict=True)     # DataParallel prefix is stripped
    assert torch.equal(m.state_dict()["hourglass.conv5.0.weight"], p["hourglass.conv5.0.weight"])
    assert pack_ssr(m.ssr_upsample).numel() == 189
    with pytest.raises(KeyError):
        DisparityHotPath(64).load_state_dict({"gamma": torch.zeros(1)}, strict=True)
    with pytest.raises(NotImplementedError):
        DisparityHotPath(32)                # 16 attention bins < k=24


def test_shard_bounds():
    from semstereo_b200.dist import shard_bounds
    for B in (1, 7, 8, 16):
        for N in (1, 2, 4, 8):
            spans = [shard_bounds(B, r, N) for r in range(N)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(N - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["SS_ROOT"])
import torch.distributed as dist
from semstereo_b200 import dist as sd
rank, local, world = sd.init_from_env("gloo")
class FakePath:            # stands in for DisparityHotPath on CPU: any per-sample function
    att_weights_only = False
    def __call__(self, f8_l, *rest):
        return {"pred_up": f8_l.sum(dim=(1,), keepdim=False) * 2.0 + rest[0].mean(dim=1)}
B = int(os.environ["SS_B"])
g = torch.Generator().manual_seed(0)
names = ("f8_l", "f8_r", "f4_l", "f4_r", "cf_l", "cf_r", "spx_pred", "pred_label")
inp = {k: torch.randn(B, 3, 4, 5, generator=g) for k in names}
full = FakePath()(*[inp[k] for k in names])["pred_up"]
got = sd.ShardedHotPath(FakePath())(inp)
assert got.shape == full.shape and torch.equal(got, full), (rank, got.shape)
got0 = sd.ShardedHotPath(FakePath(), dst=0)(inp)          # DataParallel semantics: only rank 0 receives the batch
assert (got0 is None) if rank != 0 else torch.equal(got0, full), rank
class RowPath:             # row-local stand-in: row tiles with any halo must reproduce the untiled result exactly
    att_weights_only = False
    def __call__(self, f8_l, f8_r, f4_l, f4_r, cf_l, cf_r, spx, lab):
        up = lambda t, s: t.sum(1).repeat_interleave(s, 1).repeat_interleave(s, 2)
        return {"pred_up": spx.sum(1) + up(f8_l, 8) + up(f4_r, 4) + lab.mean(1)}
H, W = 640, 128
inp2 = {k: torch.randn(1, 2, H // s, W // s, generator=g) for k, s in zip(names, (8, 8, 4, 4, 4, 4, 1, 1))}
full2 = RowPath()(*[inp2[k] for k in names])["pred_up"]
for n_tiles, halo in ((3, 128), (2, 0), (5, 256)):
    got2 = sd.TiledHotPath(RowPath(), n_tiles=n_tiles, halo=halo)(inp2)
    assert torch.equal(got2, full2), (rank, n_tiles, halo)
inp3 = {k: v.transpose(-1, -2).contiguous() for k, v in inp2.items()}          # the same test along W (column bands)
full3 = RowPath()(*[inp3[k] for k in names])["pred_up"]
got3 = sd.TiledHotPath(RowPath(), n_tiles=3, halo=128, axis="w")(inp3)
assert torch.equal(got3, full3), rank
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
'''


@pytest.mark.parametrize("B", [4, 5, 1])          # even, ragged, and batch < world (rank 1 has an empty shard)
def test_sharded_gather_world2_gloo(tmp_path, B):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, SS_ROOT=ROOT, SS_B=str(B), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29610 + B), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


def test_bench_reference_arm_small():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--height", "128", "--width", "128"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_plan_row_tiles():
    from semstereo_b200.dist import plan_row_tiles
    assert plan_row_tiles(1024, 2, 384) == [(0, 896, 0, 512), (128, 1024, 512, 1024)]
    assert plan_row_tiles(1024, 1, 384) == [(0, 1024, 0, 1024)]
    tiles = plan_row_tiles(2048, 3, 128)
    assert [t[2:] for t in tiles] == [(0, 768), (768, 1408), (1408, 2048)]           # 6 + 5 + 5 units of 128 rows
    assert all(e0 % 128 == 0 and e1 % 128 == 0 and e0 <= k0 < k1 <= e1 for e0, e1, k0, k1 in tiles)
    with pytest.raises(ValueError):
        plan_row_tiles(1000, 2)
    with pytest.raises(ValueError):
        plan_row_tiles(256, 3)


def test_torch_library_ops_registered_with_fake_shapes():
    """torch.ops.semstereo_b200.* exist and their fake (meta) implementations give the reference's output shapes; on CPU tensors the
    real implementation refuses to run (no fallback)."""
    import semstereo_b200.torch_ops  # noqa: F401  (registers)
    ops_ns = torch.ops.semstereo_b200
    l = torch.empty(2, 64, 8, 16, device="meta")
    assert tuple(ops_ns.gwc_volume(l, l, 4, 8, True, True).shape) == (2, 8, 8, 8, 16)
    assert tuple(ops_ns.gwc_volume(l, l, 4, 8, False, False).shape) == (2, 8, 4, 8, 16)
    assert tuple(ops_ns.concat_volume(l, l, 4, True).shape) == (2, 128, 8, 8, 16)
    c = torch.empty(2, 24, 8, 16, device="meta")
    assert tuple(ops_ns.regression_topk(c, c, 2).shape) == (2, 1, 8, 16)
    assert tuple(ops_ns.disparity_regression(c, 12, True).shape) == (2, 8, 16)
    assert tuple(ops_ns.context_upsample(torch.empty(2, 1, 8, 16, device="meta"), torch.empty(2, 9, 32, 64, device="meta")).shape) == (2, 32, 64)
    z = torch.zeros(1, 8, 4, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops_ns.gwc_volume(z, z, 2, 2, True, False)


def test_c_abi_argument_validation_needs_no_gpu():
    """Bad arguments are rejected by the C entry points before any CUDA call: negative code + thread-local message."""
    import ctypes
    from semstereo_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)
    assert lib.ss_conv2d_tc(0, null, 64, null, 0, null, null, null, null, 0, 1, 64, 16, 16, 0, null) == -1
    assert "null pointer" in _lib.last_error()
    assert lib.ss_conv2d_tc(0, one, 48, null, 0, one, null, null, one, 0, 1, 64, 16, 16, 0, null) == -2      # Cin not a multiple of 64
    assert "multiples of 64" in _lib.last_error()
    assert lib.ss_conv2d_tc_ntile(0, 128, 128) == 128 and lib.ss_conv2d_tc_ntile(2, 128, 6) == 16 and lib.ss_conv2d_tc_ntile(0, 100, 8) == 0
    assert lib.ss_conv3d_tc_ntile(5, 64, 32) == 32 and lib.ss_conv3d_tc_ntile(5, 128, 32) == 0 and lib.ss_conv3d_tc_ntile(4, 128, 64) == 64
    assert lib.ss_concat_stem_fused(null, null, null, null, null, null, null, null, null, 0, 1, 24, 16, 16, -16, 1, null) == -1
    assert lib.ss_ssr_upsample2(one, null, one, one, one, null, one, 1, 4, 4, 6, null) == -1
    assert lib.ss_bilinear_up2(null, null, 1, 4, 4, null) == -1
