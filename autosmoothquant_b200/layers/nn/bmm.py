"""INT8 batched matmul modules with the reference's interface (autosmoothquant/layers/nn/bmm.py:1-70),
backed by ``asq_i8bmm`` (tcgen05, one launch for the whole batch when M is a multiple of the tile height)."""
import torch

from ..._CUDA import bmm_s8t_s8n_f32t, bmm_s8t_s8n_s8t, bmm_s8t_s8n_s32t


def _as_tensor(alpha):
    return alpha if torch.is_tensor(alpha) else torch.tensor(alpha)


class BMM_S8T_S8N_S8T(torch.nn.Module):
    def __init__(self, alpha):
        super().__init__()
        self.register_buffer("a", torch.tensor(alpha))

    @torch.no_grad()
    def forward(self, a, b):
        # a: [B, M, K] int8, b: [B, N, K] int8 -> [B, M, N] int8
        return bmm_s8t_s8n_s8t(a, b, self.a.item())

    @staticmethod
    def from_scale(a_scale, b_scale, output_scale):
        mod = BMM_S8T_S8N_S8T(1.0)
        mod.a = _as_tensor(a_scale * b_scale / output_scale)
        return mod


class BMM_S8T_S8N_F32T(torch.nn.Module):
    def __init__(self, alpha):
        super().__init__()
        self.register_buffer("a", torch.tensor(alpha))

    @torch.no_grad()
    def forward(self, a, b):
        # a: [B, M, K] int8, b: [B, N, K] int8 -> [B, M, N] float32
        return bmm_s8t_s8n_f32t(a, b, self.a.item())

    @staticmethod
    def from_scale(a_scale, b_scale):
        mod = BMM_S8T_S8N_F32T(1.0)
        mod.a = _as_tensor(a_scale * b_scale)
        return mod


class BMM_S8T_S8N_S32T(torch.nn.Module):
    @torch.no_grad()
    def forward(self, a, b):
        # a: [B, M, K] int8, b: [B, N, K] int8 -> [B, M, N] int32
        return bmm_s8t_s8n_s32t(a, b)
