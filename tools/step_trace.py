#!/usr/bin/env python
"""GPU experiment: per-stream timeline of one eager TTA step (an event after every launch)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tta_depth_completion_b200 import ExternalModel_Adapt, _lib
dev = torch.device('cuda:0')
h, w, dataset, mode, lr, cap = bench.WORKLOADS['kitti']
model = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=cap, device=dev)
model._prepare_head(mode)
model.load_state_dict(bench.make_checkpoint('kitti'))
model.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
model.train()
frames = [(i.to(dev), s.to(dev)) for i, s in bench.make_frames('kitti', 1, 2, 1)]
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    for i in range(4):
        model.tta_step(frames[i % 2][0], frames[i % 2][1], lr, 1.0, 1.0, 0.1)
    eng = model._last_engine
    torch.cuda.synchronize()
    F3 = ctypes.c_float * 3
    _lib.check(_lib.lib().ptta_msgchn_trace_step(eng.handle, _lib.ptr(frames[0][0]), F3(1 / 255.0, 1 / 255.0, 1 / 255.0), F3(0, 0, 0),
                                                 _lib.ptr(frames[0][1]), cap, 1.0, 1.0, 0.1, ctypes.c_void_p(st.cuda_stream)), 'trace_step')
