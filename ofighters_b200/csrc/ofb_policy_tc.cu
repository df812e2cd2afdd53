// Tensor-core (tcgen05 / TMEM) kernels of the policy forward -- placeholder until they land.
#include "ofb_common.cuh"
#include "ofb_policy_dev.cuh"

#define NOT_YET(name) do { ofb_set_error(name ": tensor engine not built yet"); return OFB_E_STATE; } while (0)
int pol_tc_conv_pool(const ofb_policy *, int, const __nv_bfloat16 *, __nv_bfloat16 *, int, int, long long, cudaStream_t) { NOT_YET("conv_pool"); }
int pol_tc_trunk12(const ofb_policy *, const uint32_t *, __nv_bfloat16 *, int, cudaStream_t) { NOT_YET("trunk12"); }
int pol_tc_up3(const ofb_policy *, const __nv_bfloat16 *, __nv_bfloat16 *, int, cudaStream_t) { NOT_YET("up3"); }
int pol_tc_up4(const ofb_policy *, const __nv_bfloat16 *, float *, float *, int *, int, cudaStream_t) { NOT_YET("up4"); }
int pol_tc_dense1(const ofb_policy *, const __nv_bfloat16 *, float *, int, cudaStream_t) { NOT_YET("dense1"); }
