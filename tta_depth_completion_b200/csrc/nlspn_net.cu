// Second translation unit of libptta_b200.so: the general-channel tcgen05 convolution family and the channel-generic
// BatchNorm / activation kernels of the NLSPN network, behind the C ABI of include/ptta_b200.h.
#include <cuda_runtime.h>
#include <string.h>
#include "common.cuh"
#include "conv_gen.cuh"
#include "../../include/ptta_b200.h"

using namespace ptta;

extern "C" {

long long ptta_convg_packed_elems(int kind, int role, int cin0, int cin1, int cout, int has_short) {
    ConvGPlan pl;
    if (convg_make_plan(pl, kind, role, 1, 16, 16, cin0, cin1, cout, has_short)) return -1;
    return convg_packed_elems(pl);
}

int ptta_convg_pack(int kind, int role, const float* weight, const float* weight_short, int cin_w, int cout_w, int cin0, int cin1,
                    int cout, int has_short, int ident_from, void* packed, ptta_stream_t stream) {
    ConvGPlan pl;
    PTTA_TRY(convg_make_plan(pl, kind, role, 1, 16, 16, cin0, cin1, cout, has_short));
    PTTA_CHECK(weight && packed && (!has_short || weight_short), "convg_pack: null pointer");
    PTTA_CHECK(cin_w <= cin0 + cin1 && cout_w <= cout, "convg_pack: weight %d->%d larger than the stored %d->%d", cin_w, cout_w, cin0 + cin1, cout);
    return launch_convg_pack(pl, kind, role, weight, weight_short, cin_w, cout_w, ident_from, (bf16*)packed, (cudaStream_t)stream);
}

int ptta_convg_run(int kind, int role, const void* x0, const void* x1, const void* packed, const float* bias, void* out, int n, int h,
                   int w, int cin0, int cin1, int cout, int has_short, ptta_stream_t stream) {
    ConvGPlan pl;
    PTTA_TRY(convg_make_plan(pl, kind, role, n, h, w, cin0, cin1, cout, has_short));
    return launch_convg(pl, (const bf16*)x0, (const bf16*)x1, (const bf16*)packed, bias, (bf16*)out, (cudaStream_t)stream);
}

}  // extern "C"
