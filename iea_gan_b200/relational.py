"""Relational Reasoning Module: one pre-LN transformer encoder block + final
LayerNorm that attends across the 40 sensor images of an event.

Host-side mirror of the reference's RRM.py (RRM.py:10-16 scaled_dot_product,
:19-63 MultiheadAttention, :66-109 EncoderBlock, :112-133 RelationalReasoning);
the arithmetic runs in libiea_sm100.so.  qkv columns are interleaved per head
([q_h | k_h | v_h] for head h, RRM.py:49-53), dropout is 0 in every reference
instantiation and is therefore a no-op module kept only for attribute parity.
"""
import torch
import torch.nn as nn

from . import engine as E


def scaled_dot_product(q, k, v):
    """softmax(q k^T / sqrt(d)) v and the attention map; q,k,v (B,h,S,d) (RRM.py:10-16)."""
    return E.module_sdp(q, k, v)


class MultiheadAttention(nn.Module):
    def __init__(self, input_dim, embed_dim, num_heads, which_linear):
        super().__init__()
        assert embed_dim % num_heads == 0, "Embedding dimension must be 0 modulo number of heads."
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.head_dim = embed_dim // num_heads
        self.which_linear = which_linear
        self.qkv_proj = which_linear(input_dim, 3 * embed_dim)
        self.o_proj = which_linear(embed_dim, embed_dim)
        self._reset_parameters()

    def _reset_parameters(self):
        for lin in (self.qkv_proj, self.o_proj):
            nn.init.xavier_uniform_(lin.weight)
            lin.bias.data.fill_(0)

    def forward(self, x, return_attention=False):
        return E.module_mha(self, x, return_attention)


class EncoderBlock(nn.Module):
    def __init__(self, input_dim, num_heads, dim_feedforward, dropout, which_linear):
        super().__init__()
        self.which_linear = which_linear
        self.self_attn = MultiheadAttention(input_dim, input_dim, num_heads, which_linear)
        self.linear_net = nn.Sequential(
            which_linear(input_dim, dim_feedforward),
            nn.Dropout(dropout),
            nn.ReLU(inplace=True),
            which_linear(dim_feedforward, input_dim),
        )
        self.norm1 = nn.LayerNorm(input_dim)
        self.norm2 = nn.LayerNorm(input_dim)
        self.dropout = nn.Dropout(dropout)
        if dropout != 0.0:
            raise NotImplementedError("built: dropout = 0 (model.py:311,794)")

    def forward(self, x):
        return E.module_rrm([self], None, x)


class RelationalReasoning(nn.Module):
    def __init__(self, num_layers, hidden_dim, **block_args):
        super().__init__()
        self.layers = nn.ModuleList([EncoderBlock(**block_args) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(hidden_dim)

    def forward(self, x):
        return E.module_rrm(list(self.layers), self.norm, x)

    def get_attention_maps(self, x):
        maps = []
        for l in self.layers:
            _, a = l.self_attn(x, return_attention=True)  # raw x, as RRM.py:130
            maps.append(a)
            x = l(x)
        return maps
