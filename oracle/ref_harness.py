"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference: from /root/reference in the build container, else
from oracle/_ref (the verbatim, git-ignored copy of the hot path's files that `make -C oracle` makes in the
build container and that travels to the GPU box with the snapshot).  Used by oracle/make_golden.py, by the
tests that pin the oracle / the CUDA path against the reference itself (they skip when neither tree exists),
by bench.py's `--impl reference` / `cpu_baseline` legs and by scripts/ref_triton_gpu.py.  Never by the product.

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

_HERE = os.path.dirname(os.path.abspath(__file__))


def _pick_root():
    env = os.environ.get('SEA_REFERENCE_ROOT')
    if env:
        return env
    for cand in ('/root/reference', os.path.join(_HERE, '_ref')):
        if os.path.isdir(os.path.join(cand, 'src', 'models', 'perlin_attention')):
            return cand
    return '/root/reference'


REFERENCE_ROOT = _pick_root()


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


def load_reference(interpret_triton=None):
    """Returns the reference's `src.models.perlin_attention` package (imported, not copied).
    interpret_triton: True = Triton's numpy interpreter (CPU), False = compiled Triton (needs a GPU), None = interpreter
    exactly when no CUDA device is visible.  The choice is fixed by the first call of a process (Triton reads
    TRITON_INTERPRET when the kernels are decorated)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    if interpret_triton is None:
        import torch
        interpret_triton = not torch.cuda.is_available()
    if interpret_triton:
        os.environ.setdefault('TRITON_INTERPRET', '1')
    else:
        os.environ['TRITON_INTERPRET'] = '0'
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
                from triton.language.extra.cuda import libdevice
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
