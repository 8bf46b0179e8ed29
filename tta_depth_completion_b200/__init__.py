"""B200-native ProxyTTA adaptation step (drop-in for seobbro/TTA-depth-completion's TTA path).

Host side: `ExternalModel_Adapt` / `MsgChnModel_Adapt` / `OutlierRemoval` mirror the reference's
Python interface; all arithmetic runs in lib/libptta_b200.so (C ABI: include/ptta_b200.h)."""
from .external_model_adapt import ExternalModel_Adapt, MsgChnModel_Adapt, OutlierRemoval, ADAPT_LOSS_TYPE  # noqa: F401
from .engine import MsgChnEngine  # noqa: F401
from .nlspn_model_adapt import NLSPNModel_Adapt  # noqa: F401
from .nlspn_engine import NlspnEngine  # noqa: F401
from .transforms import Transforms  # noqa: F401
from . import ops  # noqa: F401

__all__ = ['ExternalModel_Adapt', 'MsgChnModel_Adapt', 'NLSPNModel_Adapt', 'NlspnEngine', 'OutlierRemoval', 'Transforms', 'MsgChnEngine', 'ops', 'ADAPT_LOSS_TYPE']
