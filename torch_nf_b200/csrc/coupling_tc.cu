// RealNVP coupling layer on tcgen05 tensor cores (placeholder until the kernel lands).
#include "common.cuh"
using namespace tnf;
extern "C" {
int tnf_tc_supported(int D, int U, int L) { (void)D; (void)U; (void)L; return 0; }
size_t tnf_tc_packed_bytes(int D, int U, int L) { (void)D; (void)U; (void)L; return 0; }
int tnf_tc_pack(const float*, void*, int, int, int, int, tnf_stream_t) {
  set_error("tnf_tc_pack: tensor-core path not built");
  return TNF_ERR_UNSUPPORTED;
}
int tnf_coupling_tc(const float*, float*, float*, const void*, int64_t, int, int, int, int, int, int, const float*,
                    const float*, double*, tnf_stream_t) {
  set_error("tnf_coupling_tc: tensor-core path not built");
  return TNF_ERR_UNSUPPORTED;
}
}
