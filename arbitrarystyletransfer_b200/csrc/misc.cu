// ABI bookkeeping entry points of libast_b200.so (no kernels).
#include "common.cuh"

extern "C" int ast_abi_version(void) { return AST_ABI_VERSION; }

// Process-wide storage format of the K4 forward activations (include/ast_b200.h): fp16 unless changed.
namespace ast {
static int g_act_format = AST_DT_F16;
int act_format() { return g_act_format; }
}  // namespace ast
extern "C" int ast_set_act_format(int dtype) {
  if (dtype != AST_DT_F16 && dtype != AST_DT_BF16) return AST_E_BADARG;
  ast::g_act_format = dtype;
  return 0;
}
extern "C" int ast_get_act_format(void) { return ast::g_act_format; }

extern "C" const char* ast_error_string(int code) {
  if (code == 0) return "success";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case AST_E_BADARG: return "ast_b200: null pointer or non-positive size";
    case AST_E_SHAPE: return "ast_b200: shape not supported by this kernel";
    case AST_E_ALIGN: return "ast_b200: pointer not aligned as required";
    case AST_E_TOOMANY: return "ast_b200: more than AST_MAX_STYLES style maps";
    case AST_E_NODRIVER: return "ast_b200: CUDA driver entry point (cuTensorMapEncodeTiled) unavailable";
    case AST_E_WORKSPACE: return "ast_b200: workspace too small";
    default: return "ast_b200: unknown error";
  }
}

extern "C" int ast_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  AST_CUDA(cudaGetDevice(&dev));
  if (sm_count) AST_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) AST_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) AST_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return 0;
}
