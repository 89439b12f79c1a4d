"""TEST INFRASTRUCTURE ONLY -- numpy (fp64) port of the ORIGINAL Performer the reference carries in-tree:
/root/reference/src/dataset/lra_benchmarks/_lra_benchmarks/models/performer/performer_attention.py (JAX; jax is not
installed in this image, so the file cannot be executed -- it is restated line by line below, each function citing the
lines it follows).  It is the only Performer specification under /root/reference; the PyPI package the hot path actually
calls (performer-pytorch==1.1.4) is not vendored.  tests/test_oracle_golden.py uses this port to pin
oracle/third_party_restated/performer_pytorch and oracle/sea_oracle.py::performer_* to that in-tree spec.

Known, documented difference between the two published implementations (not an error of either restatement):
the JAX original stabilises the normaliser with `R + 2*stab*(|R| <= stab)` (:691-692), performer-pytorch adds eps = 1e-6
to the key prefix sum instead -- both are ~1e-6-relative perturbations, so the pin is at rtol 1e-5.
"""
import numpy as np


def generalized_kernel_features(data, projection_matrix, kernel_fn=lambda x: np.maximum(x, 0.0), kernel_epsilon=0.001,
                                normalize_data=True):
    """performer_attention.py:163-198 (generalized_kernel_feature_creator): phi(x) = kernel_fn((d^-1/4 x) P^T) + eps."""
    data = np.asarray(data, dtype=np.float64)
    data_normalizer = 1.0 / np.sqrt(np.sqrt(data.shape[-1])) if normalize_data else 1.0           # :180-183
    if projection_matrix is None:
        return kernel_fn(data_normalizer * data) + kernel_epsilon                                # :184-185
    data_dash = np.einsum('...id,jd->...ij', data_normalizer * data, np.asarray(projection_matrix, dtype=np.float64))   # :187-194
    return kernel_fn(data_dash) + kernel_epsilon                                                 # :195-196


def softmax_kernel_features(data, projection_matrix, is_query, normalize_data=True, eps=0.0001):
    """performer_attention.py:50-108 (nonnegative_softmax_kernel_feature_creator); data [..., T, d], one attention axis."""
    data = np.asarray(data, dtype=np.float64)
    P = np.asarray(projection_matrix, dtype=np.float64)
    data_normalizer = 1.0 / np.sqrt(np.sqrt(data.shape[-1])) if normalize_data else 1.0           # :76-81
    ratio = 1.0 / np.sqrt(P.shape[0])                                                            # :82
    data_dash = np.einsum('...id,jd->...ij', data_normalizer * data, P)                          # :86-91
    diag_data = (np.square(data).sum(-1) / 2.0) * data_normalizer * data_normalizer              # :93-95
    diag_data = diag_data[..., None]                                                             # :96
    if is_query:
        return ratio * (np.exp(data_dash - diag_data - data_dash.max(axis=-1, keepdims=True)) + eps)          # :99-102
    return ratio * (np.exp(data_dash - diag_data - data_dash.max(axis=(-1, -2), keepdims=True)) + eps)        # :103-107


def _renormalize(W, R, numerical_stabilizer):
    """performer_attention.py:691-697."""
    R = R + 2 * numerical_stabilizer * (np.abs(R) <= numerical_stabilizer)
    return W * (1.0 / R)[..., None]


def unidirectional_attention(query_prime, key_prime, value, numerical_stabilizer=0.0):
    """performer_attention.py:432-515 (_numerator / _denominator: a scan over the attention axis carrying
    p += k (x) v and p += k) + :609-642.  query_prime / key_prime [..., T, F], value [..., T, e]."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (query_prime, key_prime, value))
    T = q.shape[-2]
    W = np.empty(q.shape[:-1] + (v.shape[-1],))
    R = np.empty(q.shape[:-1])
    p_num = np.zeros(q.shape[:-2] + (k.shape[-1], v.shape[-1]))                                   # :446 init_value
    p_den = np.zeros(q.shape[:-2] + (k.shape[-1],))                                               # :492
    for t in range(T):
        p_num += np.einsum('...m,...d->...md', k[..., t, :], v[..., t, :])                        # :440
        W[..., t, :] = np.einsum('...m,...md->...d', q[..., t, :], p_num)                         # :441
        p_den += k[..., t, :]                                                                     # :488
        R[..., t] = np.einsum('...m,...m->...', q[..., t, :], p_den)                              # :489
    return _renormalize(W, R, numerical_stabilizer)


def bidirectional_attention(query_prime, key_prime, value, numerical_stabilizer=0.0):
    """performer_attention.py:643-690: Z = K'^T V, W = Q' Z, T = K'^T 1, R = Q' T."""
    q, k, v = (np.asarray(t, dtype=np.float64) for t in (query_prime, key_prime, value))
    Z = np.einsum('...tm,...td->...md', k, v)
    W = np.einsum('...tm,...md->...td', q, Z)
    Tv = k.sum(-2)
    R = np.einsum('...tm,...m->...t', q, Tv)
    return _renormalize(W, R, numerical_stabilizer)


def fast_attention(q, k, v, projection_matrix, causal):
    """The configuration the reference constructs (attention.py:159-164): causal => generalized ReLU features + unidirectional
    prefix sums; non-causal => FAVOR+ softmax features + bidirectional sums; renormalised attention."""
    if causal:
        qp = generalized_kernel_features(q, projection_matrix)
        kp = generalized_kernel_features(k, projection_matrix)
        return unidirectional_attention(qp, kp, v)
    qp = softmax_kernel_features(q, projection_matrix, True)
    kp = softmax_kernel_features(k, projection_matrix, False)
    return bidirectional_attention(qp, kp, v)
