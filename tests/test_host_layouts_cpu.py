"""CPU checks of the host-side layouts and helpers the fused kernels rely on (no GPU, no library call)."""
import torch

from autosmoothquant_b200 import _lib as L
from autosmoothquant_b200 import harness, moe


def test_interleave_gate_up_layout():
    """include/asq.h (asq_w8a8_gateup_swiglu_q8): rows [64b, 64b+32) = gate rows [32b, 32b+32), the next 32 = up rows."""
    I, K = 96, 8
    gate = torch.arange(I * K, dtype=torch.int32).view(I, K)
    up = -gate - 1
    il = L.interleave_gate_up(gate, up)
    assert il.shape == (2 * I, K)
    for b in range(I // 32):
        assert torch.equal(il[64 * b:64 * b + 32], gate[32 * b:32 * b + 32])
        assert torch.equal(il[64 * b + 32:64 * b + 64], up[32 * b:32 * b + 32])
    v = L.interleave_gate_up(torch.arange(I), torch.arange(I) + 1000)  # vectors (scales, biases) use the same order
    assert v[:32].tolist() == list(range(32)) and v[32:64].tolist() == list(range(1000, 1032))


def test_rope_tables_blocked_index_formula():
    """Entry (pos, col) of the [S, head_dim] table lives at ((col // 8) * S + pos) * 8 + col % 8."""
    S, hd = 37, 128
    table = torch.arange(S * hd, dtype=torch.float32).view(S, hd)
    flat = L.rope_tables_blocked(table).reshape(-1)
    for pos, col in ((0, 0), (5, 7), (36, 127), (11, 64), (20, 9)):
        assert flat[((col // 8) * S + pos) * 8 + col % 8] == table[pos, col]


def test_route_tokens_layout_cpu():
    g = torch.Generator().manual_seed(0)
    E, T, k = 8, 333, 2
    sel = torch.stack([torch.randperm(E, generator=g)[:k] for _ in range(T)])
    sel = torch.where(sel == 3, torch.full_like(sel, 4), sel)  # expert 3 gets no tokens
    dest, blk, m_pad = moe.route_tokens(sel, E)
    flat = sel.reshape(-1)
    assert m_pad % 256 == 0 and blk.numel() == m_pad // 128 and blk.dtype == torch.int32
    assert dest.unique().numel() == dest.numel() and int(dest.max()) < m_pad
    assert torch.equal(blk[dest // 128].long(), flat)
    counts = torch.bincount(flat, minlength=E)
    used = int(((counts + 255) // 256 * 256).sum())
    assert (blk[used // 128:] == -1).all() and (blk[:used // 128] >= 0).all() and not (blk == 3).any()
    starts = torch.cumsum((counts + 255) // 256 * 256, 0) - (counts + 255) // 256 * 256
    assert all(int(s) % 256 == 0 for s in starts)


def test_quant_config_normalisation():
    qc = harness.normalise_quant_config({"type": "fp8"})
    assert qc["type"] == "fp8_e4m3" and qc["activation_scheme"] == "dynamic"
    assert harness.normalise_quant_config({})["qkv"] == "per-tensor"
    for bad in ({"type": "int4"}, {"qkv": "per-channel"}):
        try:
            harness.normalise_quant_config(bad)
        except ValueError:
            continue
        raise AssertionError(f"{bad} accepted")


def test_fp8_fused_columns_layout_cpu():
    """harness.FP8FusedColumnsLinear: e4m3 bytes concatenated along N, one weight scale per output column, fp32 bias."""
    from autosmoothquant_b200.layers.nn.linear import FP8LinearDynamic

    torch.manual_seed(0)
    mods = [FP8LinearDynamic.from_float(torch.nn.Linear(32, n, bias=True), act_quant="per-token", reference_compat=False)
            for n in (16, 8, 8)]
    fused = harness.FP8FusedColumnsLinear(mods)
    assert fused.sizes == [16, 8, 8] and fused.out_features == 32 and fused.weight.dtype == torch.float8_e4m3fn
    assert torch.equal(fused.weight.view(torch.uint8)[16:24], mods[1].weight.view(torch.uint8))
    assert fused.col_scale.dtype == torch.float32 and fused.col_scale.shape == (32,)
    assert torch.equal(fused.col_scale[24:], torch.full((8,), float(mods[2].weight_scale)))
    assert fused.bias.dtype == torch.float32 and torch.equal(fused.bias[:16], mods[0].bias)
    import pytest
    with pytest.raises(ValueError):  # per-tensor-dynamic modules (the converter's quirk product) are not fusable
        harness.FP8FusedColumnsLinear([FP8LinearDynamic.from_float(torch.nn.Linear(32, 8))])
    with pytest.raises(RuntimeError):  # and there is no CPU fallback for the forward
        fused(torch.randn(4, 32))


def test_route_tokens_properties_hypothesis():
    """Property test of the expert-sorted, 256-row padded layout (the grouped kernel's contract): every routed slot gets
    its own row inside its expert's segment, segments are 256-aligned and ordered by expert, slots keep their original
    order inside a segment (stable), unused 128-row blocks are marked -1, and the worst-case bound holds — for any expert
    count, top-k and routing, including experts with no tokens and a single token."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 8), st.integers(1, 3), st.integers(1, 300), st.integers(0, 2 ** 31 - 1), st.booleans())
    def check(E, k, T, seed, skewed):
        g = torch.Generator().manual_seed(seed)
        if skewed:  # most tokens to one expert, some experts empty
            sel = torch.where(torch.rand(T, k, generator=g) < 0.85, torch.zeros(T, k, dtype=torch.int64),
                              torch.randint(0, E, (T, k), generator=g))
        else:
            sel = torch.randint(0, E, (T, k), generator=g)
        dest, blk, m_pad = moe.route_tokens(sel, E)
        flat = sel.reshape(-1)
        n = flat.numel()
        assert m_pad % 256 == 0 and m_pad >= n and m_pad <= n + E * 255 and blk.numel() == m_pad // 128
        assert dest.shape == (n,) and dest.unique().numel() == n and int(dest.min()) >= 0 and int(dest.max()) < m_pad
        assert torch.equal(blk[dest // 128].long(), flat)  # every slot sits in a block of its own expert
        counts = torch.bincount(flat, minlength=E)
        padded = (counts + 255) // 256 * 256
        starts = torch.cumsum(padded, 0) - padded
        for e in range(E):
            rows = dest[flat == e]
            assert torch.equal(rows, starts[e] + torch.arange(int(counts[e])))  # contiguous from the segment start, stable
        used = int(padded.sum()) // 128
        assert (blk[used:] == -1).all() and (blk[:used] >= 0).all()
        real = blk[:used]
        assert torch.equal(real, torch.sort(real)[0])  # segments ordered by expert

    check()


def _u8(t):
    return t.view(torch.uint8) if t.dtype == torch.float8_e4m3fn else t


def test_tensor_parallel_shards_of_the_fused_projections_cpu():
    """build-time layout of the tensor-parallel stack (BASELINE configs 4 / 5), INT8 and FP8: every rank draws the same
    seeded full-size weights, keeps its Megatron shard of each projection and THEN fuses q|k|v / gate|up — so a rank's fused
    weight must be the concatenation of the r-th column shard of q, k, v (gate, up) of the unsharded model, its per-column
    scales must be the unsharded ones (per-tensor weight scales are shared by all shards), and the row-parallel
    projections must hold the r-th slice of K."""
    world = 2
    cfg = harness.TINY_GQA
    for qc in ({}, {"type": "fp8", "qkv": "per-token", "out": "per-token", "fc1": "per-token", "fc2": "per-token"}):
        full = harness.QuantDecoder(cfg, qc, device="cpu", seed=7, fuse_projections=True).layers[0]
        q_n, k_n, v_n = full.qkv_sizes
        I = cfg.intermediate
        fw, gw = _u8(full.qkv_proj.weight), _u8(full.gate_up_proj.weight)
        for r in range(world):
            shard = harness.QuantDecoder(cfg, qc, device="cpu", seed=7, fuse_projections=True, tp=(r, world)).layers[0]
            assert shard.qkv_sizes == [q_n // world, k_n // world, v_n // world]
            want = torch.cat([fw[r * q_n // world:(r + 1) * q_n // world],
                              fw[q_n + r * k_n // world:q_n + (r + 1) * k_n // world],
                              fw[q_n + k_n + r * v_n // world:q_n + k_n + (r + 1) * v_n // world]])
            assert torch.equal(_u8(shard.qkv_proj.weight), want)
            want_gu = torch.cat([gw[r * I // world:(r + 1) * I // world], gw[I + r * I // world:I + (r + 1) * I // world]])
            assert torch.equal(_u8(shard.gate_up_proj.weight), want_gu)
            if qc:  # FP8: one weight scale per output column
                fcs, scs = full.qkv_proj.col_scale, shard.qkv_proj.col_scale
                assert scs.shape == (sum(shard.qkv_sizes),)
                assert torch.equal(scs[:q_n // world], fcs[:q_n // world]) and torch.equal(scs[-1:], fcs[-1:])
                assert torch.equal(shard.gate_up_proj.col_scale[:1], full.gate_up_proj.col_scale[:1])
                assert float(shard.o_proj.weight_scale) == float(full.o_proj.weight_scale)
            else:   # INT8: the three (two) dequant scalars of the fused-W_pack module
                for name in full.qkv_proj._scale_names:
                    assert float(getattr(shard.qkv_proj, name)) == float(getattr(full.qkv_proj, name))
                assert float(shard.down_proj.dequant_scale) == float(full.down_proj.dequant_scale)
            H = cfg.hidden
            assert torch.equal(_u8(shard.o_proj.weight), _u8(full.o_proj.weight)[:, r * H // world:(r + 1) * H // world])
            assert torch.equal(_u8(shard.down_proj.weight), _u8(full.down_proj.weight)[:, r * I // world:(r + 1) * I // world])
