"""Oracle: SupportSets RBF warper (test infrastructure, CPU torch).

Follows /root/reference/lib/support_sets.py:
  * parameter shapes and initialisation            :33-79
  * forward (one-hot gather, RBF gradient, L2 norm) :81-101
"""
import math
import torch


def init_state(num_support_sets, num_support_dipoles, dim, gamma=None, generator=None,
               dtype=torch.float32):
    """State dict {SUPPORT_SETS [K, 2*D*d], ALPHAS [K, 2D], LOGGAMMA [K, 1]}.

    Antipodal dipoles on K spheres with radii arange(1, 4, 3/K) (reference :39-54),
    alphas alternating +1/-1 (:63-70), log-gamma = log(gamma) with gamma = 1/d by
    default (:26,78; train.py:158).
    """
    K, D, d = num_support_sets, num_support_dipoles, dim
    if gamma is None:
        gamma = 1.0 / d
    radii = torch.arange(1.0, 4.0, 3.0 / K)[:K]
    sv = torch.randn(K, D, d, generator=generator)
    pairs = torch.stack([sv, -sv], dim=2).reshape(K, 2 * D, d)
    pairs = radii.view(K, 1, 1) * pairs / pairs.norm(dim=2, keepdim=True)
    alphas = torch.tensor([1.0, -1.0]).repeat(D).expand(K, 2 * D).contiguous()
    return {
        'SUPPORT_SETS': pairs.reshape(K, 2 * D * d).to(dtype),
        'ALPHAS': alphas.to(dtype),
        'LOGGAMMA': torch.full((K, 1), math.log(gamma), dtype=dtype),
    }


def forward(state, mask, z, learn_gammas=True, gamma=None):
    """Unit-norm gradient of the selected warping function at z (reference :81-101).

    mask: [B, K] one-hot float, z: [B, d].  Returns [B, d].
    """
    S, A, LG = state['SUPPORT_SETS'], state['ALPHAS'], state['LOGGAMMA']
    d = z.shape[1]
    n = A.shape[1]
    sets = (mask @ S).reshape(-1, n, d)                       # :83-84
    alphas = (mask @ A).unsqueeze(2)                          # :87
    if learn_gammas:
        gammas = torch.exp(mask @ LG).unsqueeze(2)            # :91
    else:
        g = gamma if gamma is not None else 1.0 / d
        gammas = torch.full((z.shape[0], n, 1), g, dtype=z.dtype, device=z.device)   # :93
    diff = z.unsqueeze(1) - sets                              # :96
    sq = torch.norm(diff, dim=2) ** 2                         # :98 (norm()**2, as the reference)
    grad = -2.0 * (alphas * gammas * torch.exp(-gammas * sq.unsqueeze(2)) * diff).sum(dim=1)
    return grad / torch.norm(grad, dim=1, keepdim=True)       # :101, no epsilon


def one_hot(indices, num_support_sets, dtype=torch.float32):
    """The mask lib/trainer.py:227-231 builds with a Python loop."""
    m = torch.zeros(indices.shape[0], num_support_sets, dtype=dtype, device=indices.device)
    m[torch.arange(indices.shape[0], device=indices.device), indices] = 1.0
    return m
