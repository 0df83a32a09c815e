// emu_library.cpp — the WHOLE C ABI of libghicp_b200.so built for the CPU through the host emulation shim
// (tests/harness/cuda_emu): ghicp_capi.cu (context, iteration orchestration, every extern "C" entry point) on top of the
// product's own kernels of ghicp_kernels.cu / ghicp_stream.cu / ghicp_auction.cu / ghicp_fpfh.cu / ghicp_solvers.cu /
// ghicp_prep.cu, every CUDA thread a fiber.  The streaming kernel's handful of PTX wrappers (mbarrier, cp.async.bulk, packed
// f32x2 arithmetic) have host stand-ins inside ghicp_stream.cu; the tcgen05 FD build of ghicp_fdtc.cu cannot be emulated and
// is reported "not supported", which routes the FD build onto the POPC kernel exactly as GHICP_FD_POPC does on a GPU.
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
#include "../../gh-icp_b200/csrc/ghicp_stream.cu"

namespace ghicp_b200 {
// ghicp_fdtc.cu: tcgen05 / TMEM inline PTX, no emulation: the tensor-core FD build reports "not supported" -> POPC kernel.
cudaError_t launch_fd_bsc_tc(Ctx *) { return cudaErrorNotSupported; }
}  // namespace ghicp_b200

#include "../../gh-icp_b200/csrc/ghicp_capi.cu"
