import torch
from torch import nn
from oracle.performer_oracle import causal_dot_product_chunked, EPS_ATTN


class CausalLinearAttention(nn.Module):
    def __init__(self, query_dimensions, feature_map=None, eps=EPS_ATTN):
        super().__init__()
        self.feature_map = feature_map(query_dimensions)
        self.eps = eps

    def forward(self, queries, keys, values, attn_mask, query_lengths, key_lengths):
        self.feature_map.new_feature_map(queries.device)
        Q = self.feature_map.forward_queries(queries)
        K = self.feature_map.forward_keys(keys)
        if not attn_mask.lower_triangular:
            raise RuntimeError("CausalLinearAttention only supports full lower triangular masks")
        if key_lengths is not None:
            K = K * key_lengths.float_matrix[:, :, None, None]
        Z = 1 / (torch.einsum("nlhi,nlhi->nlh", Q, K.cumsum(1)) + self.eps)
        V = causal_dot_product_chunked(Q.permute(0, 2, 1, 3).contiguous(),
                                       K.permute(0, 2, 1, 3).contiguous(),
                                       values.permute(0, 2, 1, 3).contiguous()).permute(0, 2, 1, 3)
        return V * Z[:, :, :, None]


class AttentionLayer(nn.Module):
    def __init__(self, attention, d_model, n_heads, d_keys=None, d_values=None):
        super().__init__()
        d_keys = d_keys or (d_model // n_heads)
        d_values = d_values or (d_model // n_heads)
        self.inner_attention = attention
        self.query_projection = nn.Linear(d_model, d_keys * n_heads)
        self.key_projection = nn.Linear(d_model, d_keys * n_heads)
        self.value_projection = nn.Linear(d_model, d_values * n_heads)
        self.out_projection = nn.Linear(d_values * n_heads, d_model)
        self.n_heads = n_heads

    def forward(self, queries, keys, values, attn_mask, query_lengths, key_lengths):
        N, L, _ = queries.shape
        _, S, _ = keys.shape
        H = self.n_heads
        queries = self.query_projection(queries).view(N, L, H, -1)
        keys = self.key_projection(keys).view(N, S, H, -1)
        values = self.value_projection(values).view(N, S, H, -1)
        new_values = self.inner_attention(queries, keys, values, attn_mask, query_lengths,
                                          key_lengths).reshape(N, L, -1)
        return self.out_projection(new_values)
