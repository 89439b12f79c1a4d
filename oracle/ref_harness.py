"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference from /root/reference (this container
only; the path does not exist on the GPU box, so nothing under `-m gpu`, smoke() or bench.py may
import this file).  Used by oracle/make_golden.py and by the CPU tests that pin the oracle against
the reference itself (they skip when /root/reference is absent).

Recipe = SURVEY.md appendix B:
  * `src.models.hf_bert` fails to import under transformers 5.x; it is only used as the name of a
    config container (config.py:10), so it is pre-seeded with transformers.BertConfig.
  * `performer_pytorch` (third party, pinned 1.1.4, not vendored) is supplied by the restatement
    in oracle/third_party_restated/.
  * Triton kernels are run through Triton's CPU interpreter (TRITON_INTERPRET=1), with the
    one-line shim `tl.math.round = libdevice.round` the reference needs on Triton >= 3.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('SEA_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src', 'models', 'perlin_attention'))


def _patch_interpreter_boolops():
    """The Triton COMPILER lowers python `a and b` / `a or b` on tensors to logical_and / logical_or;
    the numpy interpreter evaluates them with python truthiness (a block tensor is always truthy), which
    silently drops masks such as `(j < col_len) and mask` in the reference kernels
    (causal_resize_m_to_t.py:572, flat_csr_elmul.py:80) and reads out of bounds.  Teach the
    interpreter's AST pass the compiler's meaning so the interpreted run equals the compiled one."""
    import ast
    from triton.runtime import interpreter as ti
    if getattr(ti.ASTTransformer, '_sea_boolop_patched', False):
        return

    def visit_BoolOp(self, node):
        self.generic_visit(node)
        op = ast.BitAnd() if isinstance(node.op, ast.And) else ast.BitOr()
        out = node.values[0]
        for v in node.values[1:]:
            out = ast.BinOp(left=out, op=op, right=v)
        return ast.copy_location(out, node)

    _orig_assign = ti.ASTTransformer.visit_Assign

    def visit_Assign(self, node):
        node.value = self.visit(node.value)  # the stock pass does not descend into the RHS
        return _orig_assign(self, node)

    ti.ASTTransformer.visit_BoolOp = visit_BoolOp
    ti.ASTTransformer.visit_Assign = visit_Assign
    ti.ASTTransformer._sea_boolop_patched = True


def load_reference(interpret_triton: bool = True):
    """Returns the reference's `src.models.perlin_attention` package (imported, not copied)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    if interpret_triton:
        os.environ.setdefault('TRITON_INTERPRET', '1')
    import transformers  # noqa: F401
    restated = os.path.join(_HERE, 'third_party_restated')
    for p in (REFERENCE_ROOT, restated):
        if p not in sys.path:
            sys.path.insert(0, p)
    if 'src.models.hf_bert' not in sys.modules:
        stub = types.ModuleType('src.models.hf_bert')
        stub.BertConfig = transformers.BertConfig
        sys.modules['src.models.hf_bert'] = stub
    try:
        import triton.language as tl
        if os.environ.get('TRITON_INTERPRET', '0') == '1':
            _patch_interpreter_boolops()
        if not hasattr(tl.math, 'round'):
            if os.environ.get('TRITON_INTERPRET', '0') == '1':
                # libdevice externs do not exist in the numpy interpreter: exact round-half-away-
                # from-zero (what libdevice's roundf computes), done in fp64 so x+0.5 is exact.
                def _round_half_away(x):
                    x64 = x.to(tl.float64)
                    r = tl.floor(tl.abs(x64) + 0.5)
                    return tl.where(x64 < 0, -r, r).to(tl.float32)
                tl.math.round = _round_half_away
            else:
                from triton.language.extra import libdevice
                tl.math.round = libdevice.round
    except Exception:  # triton missing: dense path still works
        pass
    import src.models.perlin_attention as pa  # noqa: E402
    return pa


def build_reference_attention(H, d, T_max, k, P, nbf, causal, k_flatten_dim=None, seed=42, **pcfg_kw):
    """Constructs the reference PerlinAttention exactly like test_perlin_opt_causality.py:110-173."""
    import torch
    import transformers
    pa = load_reference()
    from src.utils import seed as ref_seed
    ref_seed(seed)
    cfg = transformers.BertConfig(hidden_size=H * d, num_attention_heads=H, max_position_embeddings=T_max)
    if k_flatten_dim is None:
        k_flatten_dim = 'causal_batch' if causal else 'batch'
    pcfg = pa.PerlinAttentionConfig(performer_nb_factor=nbf, k=k, attention_predictor_length=P, causal=causal,
                                    k_flatten_dim=k_flatten_dim, context_output_method='mix', **pcfg_kw)
    m = pa.PerlinAttention(cfg, pcfg).eval()
    return m


def causal_additive_mask(T, dtype):
    import torch
    fp_min = torch.finfo(torch.float16).min / 2 if dtype in (torch.float16, torch.bfloat16) else torch.finfo(torch.float32).min / 2
    m = (torch.arange(T).view(1, T) > torch.arange(T).view(T, 1)) * fp_min
    return m.view(1, 1, T, T).to(dtype)
