"""Test scaffolding: a stand-in for the reference WanModel *weight container* (attribute and parameter
names of animation/diffsynth/models/wan_video_dit.py:271-336, no forward) so the drop-in boundary can be
exercised on the GPU box where /root/reference does not exist."""
import torch
import torch.nn as nn


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q, self.k, self.v, self.o = (nn.Linear(d, d) for _ in range(4))
        self.norm_q, self.norm_k = _Rms(d), _Rms(d)


class _Rms(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))


class _Block(nn.Module):
    def __init__(self, d, f, heads, eps):
        super().__init__()
        self.dim, self.num_heads, self.ffn_dim = d, heads, f
        self.self_attn, self.cross_attn = _Attn(d), _Attn(d)
        self.norm1 = nn.LayerNorm(d, eps=eps, elementwise_affine=False)
        self.norm2 = nn.LayerNorm(d, eps=eps, elementwise_affine=False)
        self.norm3 = nn.LayerNorm(d, eps=eps)
        self.ffn = nn.Sequential(nn.Linear(d, f), nn.GELU(approximate="tanh"), nn.Linear(f, d))
        self.modulation = nn.Parameter(torch.zeros(1, 6, d))


class _Head(nn.Module):
    def __init__(self, d, out):
        super().__init__()
        self.head = nn.Linear(d, out)
        self.modulation = nn.Parameter(torch.zeros(1, 2, d))


class WanContainer(nn.Module):
    require_vae_embedding = False
    require_clip_embedding = False

    def __init__(self, cfg):
        super().__init__()
        d = cfg.dim
        self.dim, self.in_dim, self.freq_dim, self.patch_size = d, cfg.in_dim, cfg.freq_dim, tuple(cfg.patch_size)
        self.seperated_timestep = True
        self.patch_embedding = nn.Conv3d(cfg.in_dim, d, kernel_size=cfg.patch_size, stride=cfg.patch_size)
        self.text_embedding = nn.Sequential(nn.Linear(cfg.text_dim, d), nn.GELU(approximate="tanh"), nn.Linear(d, d))
        self.time_embedding = nn.Sequential(nn.Linear(cfg.freq_dim, d), nn.SiLU(), nn.Linear(d, d))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(d, 6 * d))
        self.blocks = nn.ModuleList([_Block(d, cfg.ffn_dim, cfg.num_heads, cfg.eps) for _ in range(cfg.num_layers)])
        self.head = _Head(d, cfg.out_dim * 4)
        for p in self.parameters():
            p.requires_grad_(False)
