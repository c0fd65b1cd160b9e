"""Drop-in alias: ``import kfac`` / ``import kfac.backend as backend`` resolve to
the B200-native package, so the reference's training scripts run unchanged."""
import sys as _sys

import kfac_pytorch_b200 as _impl
from kfac_pytorch_b200 import *  # noqa: F401,F403
from kfac_pytorch_b200 import backend, kfac_mappers, get_kfac_module, DP_KFAC, KFACParamScheduler  # noqa: F401

_sys.modules[__name__ + '.backend'] = backend
