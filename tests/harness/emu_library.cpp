// emu_library.cpp — the WHOLE C ABI of libghicp_b200.so built for the CPU through the host emulation shim
// (tests/harness/cuda_emu): ghicp_capi.cu (context, iteration orchestration, every extern "C" entry point) on top of the
// product's own kernels of ghicp_kernels.cu / ghicp_auction.cu / ghicp_fpfh.cu / ghicp_solvers.cu / ghicp_prep.cu, every CUDA
// thread a fiber.  What cannot be emulated — the inline-PTX kernels of ghicp_stream.cu (TMA ring) and ghicp_fdtc.cu
// (tcgen05) — is reported "not supported", which routes the ABI onto its all-double kernels exactly as
// ghicp_config.force_exact / GHICP_FD_POPC do on a GPU.
// TEST INFRASTRUCTURE ONLY: tests/test_emulated_abi.py loads it explicitly to run the Python and C++ host layers, the
// command-line driver and the GPU test functions themselves on a machine without a GPU.  The product never loads it;
// libghicp_b200.so has no CPU path.
#define GHICP_EMU_HOST 1
#define GHICP_EMU_LIBRARY 1
#include "../../gh-icp_b200/csrc/ghicp_kernels.cu"
#include "../../gh-icp_b200/csrc/ghicp_auction.cu"
#include "../../gh-icp_b200/csrc/ghicp_fpfh.cu"
#include "../../gh-icp_b200/csrc/ghicp_solvers.cu"
#include "../../gh-icp_b200/csrc/ghicp_prep.cu"
#include "../../gh-icp_b200/csrc/ghicp_comm.cu"

namespace ghicp_b200 {
// ghicp_stream.cu / ghicp_fdtc.cu: inline PTX, no emulation.  The ABI never reaches the stream launchers when the context
// is created with use_fast = false (forced below); the tensor-core FD build reports "not supported" -> POPC kernel.
cudaError_t launch_fd_bsc_tc(Ctx *) { return cudaErrorNotSupported; }
cudaError_t launch_stream_prep(Ctx *, const CostParams &, int) { return cudaErrorNotSupported; }
cudaError_t launch_stream_seed(Ctx *, const CostParams &, bool) { return cudaErrorNotSupported; }
cudaError_t launch_stream(Ctx *, const CostParams &, int, bool) { return cudaErrorNotSupported; }
cudaError_t launch_stream_resolve(Ctx *, const CostParams &, bool) { return cudaErrorNotSupported; }
cudaError_t launch_stream_gate(Ctx *, const CostParams &) { return cudaErrorNotSupported; }
cudaError_t launch_colmerge(Ctx *) { return cudaErrorNotSupported; }   // multi-GPU column merge (sharded runs only)
int stream_num_parts(const Ctx *) { return 1; }
}  // namespace ghicp_b200

#include "../../gh-icp_b200/csrc/ghicp_capi.cu"
