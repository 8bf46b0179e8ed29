"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference (seobbro/TTA-depth-completion).

Imports the unmodified reference from /root/reference (read-only) with the mechanical shims of
SURVEY.md Appendix C so that its own code can be executed on a CPU-only box.  Used by
`oracle/gen_golden.py` to produce the committed fixtures under tests/golden/, and by the
"not gpu" tests when /root/reference happens to be present.  Nothing in the product package
(`tta-depth-completion_b200/`) may import this module; /root/reference does not exist on the GPU
box, so nothing that runs there may need it either.

Shims (none of them touches arithmetic):
  1. src/msg_chn_model_adapt.py:1 starts with the 26-char garbage `src/msg_chn_model_adapt.py`
     glued to `import torch` -> SyntaxError as shipped; we exec the text minus that prefix.
  2. `.cuda()` / `.to('cuda')` are hard-coded in the networks
     (network_exp_msg_chn_adapt.py:31-33,511,1034-1036) -> made no-ops on a CPU-only host.
  3. matplotlib (src/log_utils.py:24) is not installed -> stub module.
"""
import os
import sys
import types
import importlib

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('PTTA_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src'))


_loaded = {}


def _neutralise_cuda():
    if torch.cuda.is_available():
        return
    nn.Module.cuda = lambda self, device=None: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    orig_to = nn.Module.to

    def to(self, *args, **kwargs):
        def fix(a):
            if isinstance(a, torch.device) and a.type == 'cuda':
                return torch.device('cpu')
            if isinstance(a, str) and a.startswith('cuda'):
                return 'cpu'
            return a
        args = tuple(fix(a) for a in args)
        kwargs = {k: fix(v) for k, v in kwargs.items()}
        return orig_to(self, *args, **kwargs)
    if not getattr(nn.Module.to, '_ptta_shim', False):
        to._ptta_shim = True
        nn.Module.to = to


def enable_cpu_syncbn():
    """SyncBatchNorm.forward refuses CPU tensors; at world size 1 it is F.batch_norm on the local batch (torch/nn/modules/batchnorm.py:
    `need_sync` is False without a process group), which is what this replacement calls -- with the module's own buffers, so the
    `running_mean = None` the reference sets in adapt_parameters('meta_bn') (src/nlspn_model_adapt.py:329-333) selects batch statistics
    in train AND eval mode exactly as on the GPU.  The driver calls convert_syncbn() before adapt_parameters (src/tta_main.py:327-339),
    which also turns the heads' BatchNorm1d layers into SyncBatchNorm instances."""
    import torch.nn.functional as F

    def forward(self, x):
        if self.momentum is None:
            eaf = 0.0
        else:
            eaf = self.momentum
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                eaf = 1.0 / float(self.num_batches_tracked)
        bn_training = self.training or (self.running_mean is None and self.running_var is None)
        rm = self.running_mean if not self.training or self.track_running_stats else None
        rv = self.running_var if not self.training or self.track_running_stats else None
        return F.batch_norm(x, rm, rv, self.weight, self.bias, bn_training, eaf, self.eps)
    if not getattr(nn.SyncBatchNorm.forward, '_ptta_shim', False):
        forward._ptta_shim = True
        nn.SyncBatchNorm.forward = forward


def load_reference():
    """Returns a namespace with ExternalModel_Adapt, OutlierRemoval, loss_utils, eval_utils."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    _neutralise_cuda()
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)           # the wrappers insert *relative* sys.path entries
    try:
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'src'))
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'external_src', 'MSG_CHN'))
        sys.path.insert(0, os.path.join(REFERENCE_ROOT, 'external_src', 'MSG_CHN', 'workspace', 'exp_msg_chn'))
        for name in ('matplotlib', 'matplotlib.pyplot'):
            if name not in sys.modules:
                try:
                    importlib.import_module(name)
                except Exception:
                    m = types.ModuleType(name)
                    m.cm = types.SimpleNamespace(get_cmap=lambda *a, **k: None)
                    sys.modules[name] = m
        if 'matplotlib' in sys.modules and 'matplotlib.pyplot' in sys.modules:
            sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
        # shim 1: exec the wrapper minus the corrupt prefix
        path = os.path.join(REFERENCE_ROOT, 'src', 'msg_chn_model_adapt.py')
        text = open(path).read()
        prefix = 'src/msg_chn_model_adapt.py'
        if text.startswith(prefix):
            text = text[len(prefix):]
        mod = types.ModuleType('msg_chn_model_adapt')
        mod.__file__ = path
        sys.modules['msg_chn_model_adapt'] = mod
        exec(compile(text, path, 'exec'), mod.__dict__)
        ema = importlib.import_module('external_model_adapt')
        net_utils = importlib.import_module('net_utils')
        loss_utils = importlib.import_module('loss_utils')
        eval_utils = importlib.import_module('eval_utils')
    finally:
        os.chdir(cwd)
    _loaded.update(ExternalModel_Adapt=ema.ExternalModel_Adapt,
                   OutlierRemoval=net_utils.OutlierRemoval,
                   loss_utils=loss_utils, eval_utils=eval_utils,
                   msg_chn_model_adapt=mod)
    return types.SimpleNamespace(**_loaded)


def build_reference_msgchn(prepare_mode, max_input_depth, max_predict_depth=100.0, min_predict_depth=0.0):
    """The driver's construction sequence, src/tta_main.py:309-346 (DDP / SyncBN skipped: world size 1)."""
    ref = load_reference()
    cwd = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            model = ref.ExternalModel_Adapt(
                model_name='msg_chn',
                max_input_depth=max_input_depth,
                min_predict_depth=min_predict_depth,
                max_predict_depth=max_predict_depth,
                device=torch.device('cuda' if torch.cuda.is_available() else 'cpu'),
                from_scratch=False, dataset_name='', offset=True)
            model._prepare_head(prepare_mode)
    finally:
        os.chdir(cwd)
    return model
