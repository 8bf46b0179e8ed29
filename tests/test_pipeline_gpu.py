"""GPU integration: the protocol from PNG FILES to an adaptation step -- host PNG decode (csrc/png_host.cu) -> uint8 / uint16 H2D ->
`ops.input_stage` (float conversion, depth / 256, validity, bottom crop; src/data_utils.py:134-234, src/datasets.py:83-170) ->
`ExternalModel_Adapt.tta_step` -- gives bit for bit the losses and adapted weights of the step fed with the reference loaders' fp32
tensors (PIL + numpy, as src/data_utils.py computes them)."""
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PIL = pytest.importorskip('PIL.Image')
from oracle import msgchn_oracle as O

DEV = 'cuda'


def test_png_files_to_adapted_weights():
    from tta_depth_completion_b200 import ExternalModel_Adapt, ops
    mode, cap, lr = 'meta_selfsup_seq_2layers_ema', 80.0, 1e-4
    h0, w0, h, w = 75, 150, 64, 128
    sd = O.make_synthetic_checkpoint(0, mode)

    def model():
        m = ExternalModel_Adapt('msg_chn', 0.0, 100.0, max_input_depth=cap, device=torch.device(DEV))
        m._prepare_head(mode)
        m.load_state_dict(sd)
        m.set_image_normalization((1 / 255.0,) * 3, (0.0,) * 3)
        m.train()
        return m
    a, b = model(), model()
    for t in range(2):
        image, sparse, _ = O.synthetic_frame(3, t, 1, h0, w0, 'kitti')
        rgb = image[0].permute(1, 2, 0).clamp(0, 255).to(torch.uint8).numpy()
        d16 = (sparse[0, 0] * 256.0).clamp(0, 65535).to(torch.int32).numpy().astype(np.uint16)
        f_rgb, f_d = io.BytesIO(), io.BytesIO()
        PIL.fromarray(rgb).save(f_rgb, format='PNG')
        PIL.fromarray(d16).save(f_d, format='PNG')
        # (1) the reference's loaders: PIL -> float32, depth / 256, crop at the bottom, horizontally centred
        img_ref = np.asarray(PIL.open(io.BytesIO(f_rgb.getvalue())).convert('RGB'), np.float32)
        z_ref = np.array(PIL.open(io.BytesIO(f_d.getvalue())), dtype=np.float32) / 256.0
        z_ref[z_ref <= 0] = 0.0
        y0, x0 = h0 - h, (w0 - w) // 2
        img_t = torch.from_numpy(img_ref[y0:y0 + h, x0:x0 + w].transpose(2, 0, 1).copy())[None].to(DEV)
        z_t = torch.from_numpy(z_ref[y0:y0 + h, x0:x0 + w].copy())[None, None].to(DEV)
        a.tta_step(img_t, z_t, lr, 1.0, 1.0, 0.1)
        # (2) the library's input path
        pin_rgb = torch.empty((1, h0, w0, 3), dtype=torch.uint8).pin_memory()
        pin_d = torch.empty((1, h0, w0), dtype=torch.uint16).pin_memory()
        ops.decode_png_rgb8(f_rgb.getvalue(), out=pin_rgb[0])
        ops.decode_png_gray16(f_d.getvalue(), out=pin_d[0])
        img_n, z_n, v_n = ops.input_stage(pin_rgb.to(DEV, non_blocking=True), pin_d.to(DEV, non_blocking=True), crop_shape=(h, w), crop_type=('bottom',))
        assert torch.equal(img_n, img_t) and torch.equal(z_n, z_t)
        assert torch.equal(v_n, (z_t > 0).float())
        b.tta_step(img_n, z_n, lr, 1.0, 1.0, 0.1)
        assert a.last_losses() == b.last_losses(), t
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
