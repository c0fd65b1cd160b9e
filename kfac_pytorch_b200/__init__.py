"""kfac_pytorch_b200 -- B200-native (sm_100a) DP-KFAC gradient preconditioner.

Drop-in for the ``kfac`` package of lzhangbv/kfac_pytorch: same ``get_kfac_module`` /
``kfac_mappers`` (``inverse``, ``eigen``, ``inverse_dp``, ``eigen_dp``) / ``DP_KFAC`` /
``KFACParamScheduler`` / ``backend`` surface (kfac/__init__.py:1-16, kfac/dp_kfac.py:4-40),
CUDA kernels underneath.  The distributed-preconditioning variants are the hot path this
package is built and benchmarked for.
"""
from . import backend
from .preconditioner import KFAC_EIGEN, KFAC_EIGEN_DP, KFAC_INV, KFAC_INV_DP
from .scheduler import KFACParamScheduler

__version__ = '0.1.0'

kfac_mappers = {
    'inverse': KFAC_INV,
    'eigen': KFAC_EIGEN,
    'inverse_dp': KFAC_INV_DP,
    'eigen_dp': KFAC_EIGEN_DP,
}


def get_kfac_module(kfac='eigen_dp'):
    return kfac_mappers[kfac]


def DP_KFAC(model, inv_type='eigen', lr=0.1, damping=0.001, fac_update_freq=1,
            kfac_update_freq=1, kl_clip=0.001, factor_decay=0.95,
            exclude_vocabulary_size=None, hook_enabled=True, exclude_parts='', **kwargs):
    """``inv_type='eigen'`` -> eigen_dp, anything else -> inverse_dp (kfac/dp_kfac.py:18-40)."""
    cls = KFAC_EIGEN_DP if inv_type == 'eigen' else KFAC_INV_DP
    return cls(model=model, lr=lr, damping=damping, fac_update_freq=fac_update_freq,
               kfac_update_freq=kfac_update_freq, kl_clip=kl_clip, factor_decay=factor_decay,
               exclude_vocabulary_size=exclude_vocabulary_size, hook_enabled=hook_enabled,
               exclude_parts=exclude_parts, **kwargs)


__all__ = ['backend', 'kfac_mappers', 'get_kfac_module', 'DP_KFAC', 'KFACParamScheduler',
           'KFAC_EIGEN_DP', 'KFAC_INV_DP', 'KFAC_EIGEN', 'KFAC_INV']
