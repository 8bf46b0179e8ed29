#!/usr/bin/env python
"""TEST INFRASTRUCTURE (checker only).  Builds the reference's own modulated-deformable-convolution CUDA kernels
(/root/reference/external_src/NLSPN/src/model/deformconv/src/cuda/modulated_deform_conv_cuda.cu + modulated_deform_im2col_cuda.cuh)
into oracle/_ref/dcn_ref*.so for sm_100a, so that the GPU parity tests can compare the B200 propagation kernels with the
reference's kernels on the same inputs.

Nothing is copied into the repository: the two source files are read where they lie, patched IN A SCRATCH DIRECTORY for
torch 2.x (`THC/THCAtomics.cuh` no longer exists, `Tensor.type()` dispatch and `.data<T>()` are
removed APIs) and compiled together with a 20-line pybind binding written here.  Only the two `modulated_*` entry
points are bound (DCNv1 and PS-RoI pooling are never called by the TTA path).  Runs only where /root/reference exists."""
import os
import re
import shutil
import sys
import tempfile

REF = '/root/reference/external_src/NLSPN/src/model/deformconv/src'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
NAME = 'dcn_ref'

BINDING = r'''
#include <torch/extension.h>
#include <vector>
at::Tensor modulated_deform_conv_cuda_forward(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&,
    const int, const int, const int, const int, const int, const int, const int, const int, const int, const int, const int);
std::vector<at::Tensor> modulated_deform_conv_cuda_backward(const at::Tensor&, const at::Tensor&, const at::Tensor&, const at::Tensor&,
    const at::Tensor&, const at::Tensor&, const int, const int, const int, const int, const int, const int, const int, const int, const int,
    const int, const int);
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("modulated_deform_conv_forward", &modulated_deform_conv_cuda_forward, "reference DCNv2 forward (CUDA)");
  m.def("modulated_deform_conv_backward", &modulated_deform_conv_cuda_backward, "reference DCNv2 backward (CUDA)");
}
'''


def patch(text):
    text = text.replace('#include <THC/THCAtomics.cuh>', '')
    text = re.sub(r'AT_DISPATCH_FLOATING_TYPES\((\w+)\.type\(\)', r'AT_DISPATCH_FLOATING_TYPES(\1.scalar_type()', text)
    text = re.sub(r'\.data<scalar_t>\(\)', '.data_ptr<scalar_t>()', text)
    text = re.sub(r'(\w+)\.type\(\)\.is_cuda\(\)', r'\1.is_cuda()', text)
    text = text.replace('AT_ASSERTM', 'TORCH_CHECK')
    return text


def build(verbose=True):
    if not os.path.isdir(REF):
        raise RuntimeError('%s not found: the reference kernels can only be built where the reference is mounted' % REF)
    os.makedirs(OUT, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix='dcn_ref_')
    try:
        os.makedirs(os.path.join(scratch, 'cuda'))
        for f in ('modulated_deform_conv_cuda.cu', 'modulated_deform_im2col_cuda.cuh', 'modulated_deform_conv_cuda.h'):
            with open(os.path.join(REF, 'cuda', f)) as fh:
                text = patch(fh.read())
            with open(os.path.join(scratch, 'cuda', f), 'w') as fh:      # the .cu includes "cuda/...cuh"
                fh.write(text)
        with open(os.path.join(scratch, 'binding.cpp'), 'w') as fh:
            fh.write(BINDING)
        os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0a')
        from torch.utils.cpp_extension import load
        build_dir = os.path.join(scratch, 'build')
        os.makedirs(build_dir)
        load(name=NAME, sources=[os.path.join(scratch, 'binding.cpp'), os.path.join(scratch, 'cuda', 'modulated_deform_conv_cuda.cu')],
             extra_include_paths=[scratch], extra_cuda_cflags=['-gencode', 'arch=compute_100a,code=sm_100a', '-O3'],
             build_directory=build_dir, verbose=verbose, is_python_module=False)
        so = [f for f in os.listdir(build_dir) if f.endswith('.so')]
        if not so:
            raise RuntimeError('build produced no shared object')
        shutil.copy(os.path.join(build_dir, so[0]), os.path.join(OUT, NAME + '.so'))
        return os.path.join(OUT, NAME + '.so')
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == '__main__':
    print(build())
