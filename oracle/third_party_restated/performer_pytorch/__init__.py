"""TEST INFRASTRUCTURE ONLY -- restatement of the third-party `performer-pytorch` FastAttention.

The reference imports `performer_pytorch.FastAttention` (attention.py:13, 159-164;
attention_state.py:13; common/performer.py:3) from the PyPI package pinned
`performer-pytorch==1.1.4` (environment.yml:129). That package is NOT vendored under
/root/reference and cannot be installed here (no network), so its published algorithm
(Choromanski et al., "Rethinking Attention with Performers", FAVOR+) is restated below.
Cross-checked against the in-tree JAX original the reference carries under
src/dataset/lra_benchmarks/_lra_benchmarks/models/performer/performer_attention.py:
  :50-108  softmax (FAVOR+) features,
  :163-198 generalized (ReLU) features incl. kernel_epsilon and d^-1/4 normalisation,
  :336-377 gaussian orthogonal random matrix,
  :432-515 causal prefix-sum numerator / denominator,
and against the in-tree torch restatement attention_state.py:80-98 (chunked causal form).

Parity status: **unpinned** -- no reference test holds a golden vector at this boundary
(SURVEY.md section 8c). The product and this oracle both take the projection matrix as an
explicit buffer, so results never depend on the RNG used to draw it.

Only `tests/`, `__graft_entry__.smoke()`, `bench.py`'s cpu_baseline / `--impl reference`
legs and `oracle/make_golden.py` may import this module.
"""
import math
from functools import partial

import torch
from torch import nn


def softmax_kernel(data, *, projection_matrix, is_query, normalize_data=True, eps=1e-4, device=None):
    # FAVOR+ positive random features (JAX original :50-108).
    data_normalizer = (data.shape[-1] ** -0.25) if normalize_data else 1.0
    ratio = projection_matrix.shape[0] ** -0.5
    projection = projection_matrix.type_as(data)
    data_dash = torch.einsum('...id,jd->...ij', (data_normalizer * data), projection)
    diag_data = (data ** 2).sum(dim=-1)
    diag_data = (diag_data / 2.0) * (data_normalizer ** 2)
    diag_data = diag_data.unsqueeze(dim=-1)
    if is_query:
        data_dash = ratio * (torch.exp(data_dash - diag_data - torch.amax(data_dash, dim=-1, keepdim=True).detach()) + eps)
    else:
        data_dash = ratio * (torch.exp(data_dash - diag_data - torch.amax(data_dash, dim=(-1, -2), keepdim=True).detach()) + eps)
    return data_dash.type_as(data)


def generalized_kernel(data, *, projection_matrix, kernel_fn=nn.ReLU(), kernel_epsilon=0.001, normalize_data=True, device=None):
    # Generalized attention features (JAX original :163-198).
    data_normalizer = (data.shape[-1] ** -0.25) if normalize_data else 1.0
    if projection_matrix is None:
        return kernel_fn(data_normalizer * data) + kernel_epsilon
    projection = projection_matrix.type_as(data)
    data_dash = torch.einsum('...id,jd->...ij', (data_normalizer * data), projection)
    data_prime = kernel_fn(data_dash) + kernel_epsilon
    return data_prime.type_as(data)


def orthogonal_matrix_chunk(cols, device=None):
    unstructured_block = torch.randn((cols, cols), device=device)
    q, r = torch.linalg.qr(unstructured_block.cpu(), mode='reduced')
    q, r = map(lambda t: t.to(device), (q, r))
    return q.t()


def gaussian_orthogonal_random_matrix(nb_rows, nb_columns, scaling=0, device=None):
    # JAX original :336-377.
    nb_full_blocks = int(nb_rows / nb_columns)
    block_list = []
    for _ in range(nb_full_blocks):
        block_list.append(orthogonal_matrix_chunk(nb_columns, device=device))
    remaining_rows = nb_rows - nb_full_blocks * nb_columns
    if remaining_rows > 0:
        q = orthogonal_matrix_chunk(nb_columns, device=device)
        block_list.append(q[:remaining_rows])
    final_matrix = torch.cat(block_list)
    if scaling == 0:
        multiplier = torch.randn((nb_rows, nb_columns), device=device).norm(dim=1)
    elif scaling == 1:
        multiplier = math.sqrt(float(nb_columns)) * torch.ones((nb_rows,), device=device)
    else:
        raise ValueError(f'Invalid scaling {scaling}')
    return torch.diag(multiplier) @ final_matrix


def linear_attention(q, k, v):
    # Non-causal: out = q' (k'^T v) / (q' . sum_T k').
    k_cumsum = k.sum(dim=-2)
    D_inv = 1.0 / torch.einsum('...nd,...d->...n', q, k_cumsum.type_as(q))
    context = torch.einsum('...nd,...ne->...de', k, v)
    return torch.einsum('...de,...nd,...n->...ne', context, q, D_inv)


def causal_linear_attention_noncuda(q, k, v, chunk_size=128, eps=1e-6):
    # Prefix-sum form (JAX original :432-515; in-tree twin attention_state.py:80-98).
    last_k_cumsum = 0
    last_context_cumsum = 0
    outs = []
    for q, k, v in zip(*map(lambda t: t.chunk(chunk_size, dim=-2), (q, k, v))):
        k_cumsum = last_k_cumsum + k.cumsum(dim=-2)
        D_inv = 1.0 / torch.einsum('...nd,...nd->...n', q, k_cumsum.type_as(q) + eps)
        context = torch.einsum('...nd,...ne->...nde', k, v)
        context_cumsum = last_context_cumsum + context.cumsum(dim=-3)
        out = torch.einsum('...nde,...nd,...n->...ne', context_cumsum, q, D_inv)
        last_k_cumsum = k_cumsum[:, :, -1:]
        last_context_cumsum = context_cumsum[:, :, -1:]
        outs.append(out)
    return torch.cat(outs, dim=-2)


class FastAttention(nn.Module):
    def __init__(self, dim_heads, nb_features=None, ortho_scaling=0, causal=False,
                 generalized_attention=False, kernel_fn=nn.ReLU(), no_projection=False):
        super().__init__()
        nb_features = nb_features if nb_features is not None else int(dim_heads * math.log(dim_heads))
        self.dim_heads = dim_heads
        self.nb_features = nb_features
        self.ortho_scaling = ortho_scaling
        self.create_projection = partial(gaussian_orthogonal_random_matrix, nb_rows=self.nb_features,
                                         nb_columns=dim_heads, scaling=ortho_scaling)
        self.register_buffer('projection_matrix', self.create_projection())
        self.generalized_attention = generalized_attention
        self.kernel_fn = kernel_fn
        self.no_projection = no_projection
        self.causal = causal
        if causal:
            # fast_transformers' causal_product CUDA extension is absent here, exactly as on any
            # box without that build; the library then uses this chunked torch form.
            self.causal_linear_fn = causal_linear_attention_noncuda

    @torch.no_grad()
    def redraw_projection_matrix(self, device):
        projections = self.create_projection(device=device)
        self.projection_matrix.copy_(projections)
        del projections

    def forward(self, q, k, v):
        device = q.device
        if self.no_projection:
            q = q.softmax(dim=-1)
            k = torch.exp(k) if self.causal else k.softmax(dim=-2)
        elif self.generalized_attention:
            create_kernel = partial(generalized_kernel, kernel_fn=self.kernel_fn,
                                    projection_matrix=self.projection_matrix, device=device)
            q, k = map(create_kernel, (q, k))
        else:
            create_kernel = partial(softmax_kernel, projection_matrix=self.projection_matrix, device=device)
            q = create_kernel(q, is_query=True)
            k = create_kernel(k, is_query=False)
        attn_fn = linear_attention if not self.causal else self.causal_linear_fn
        return attn_fn(q, k, v)
