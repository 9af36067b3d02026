// libgkrb200ec's OWN host driver (gkr-mimc_b200/csrc/ec/ec.cu, unmodified) compiled for the CPU -- TEST INFRASTRUCTURE, never shipped
// and never loaded by the product package.
//
// ec.cu is included as is; <cuda_runtime.h> resolves to tests/emu/shim/cuda_runtime.h (host memory with poison and guard bands,
// immediate streams), and the executor -- the one piece of ec.cu that needs nvcc -- is replaced through its test seam by one that
// runs every "launch" as a loop over thread indices (in descending order when EC_HOSTBUILD_REVERSE is set in the environment).
// The result exports the same C ABI as libgkrb200ec.so, so tests/test_ec_driver_cpu.py runs the DEVICE parity tests' own bodies
// (tests/test_zz*_gpu.py) against it: staging, grow-only workspaces, base slots, error paths, statistics and the Groth16
// sequencing of the real driver are exercised without a GPU.  What it cannot show: the CUDA runtime's own behaviour, launch
// configuration, the inline-PTX carry chains (fr_device.cuh; covered by the GKR GPU tests), and speed.
#include <cuda_runtime.h>

#include <cstdlib>

#define GKRB200EC_TEST_EXECUTOR 1
namespace {
struct CudaExec {
    cudaStream_t st;
    cudaError_t err = cudaSuccess;
    template <class K, class... A>
    int launch(size_t n, A... a) {
        if (n == 0) return 0;
        static const bool reverse = getenv("EC_HOSTBUILD_REVERSE") != nullptr;
        if (reverse)
            for (size_t i = n; i-- > 0;) K::run(i, a...);
        else
            for (size_t i = 0; i < n; i++) K::run(i, a...);
        fakecuda::clock_ms() += 0.001;  // events measure "launches"
        return 1;
    }
    void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
};
}  // namespace

#include "../../gkr-mimc_b200/csrc/ec/ec.cu"
