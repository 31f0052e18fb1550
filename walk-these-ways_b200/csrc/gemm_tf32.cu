// gemm_tf32.cu — tcgen05 TF32 GEMM (placeholder until the tensor-core kernel lands; impl=1 fails loudly).
#include <cuda_runtime.h>
extern int go1_set_error(const char* m);
extern "C" int go1_gemm_tf32(int, int, int, int, int, const float*, int, const float*, int, float*, int, const float*, int, int, cudaStream_t) {
    return go1_set_error("go1_gemm impl=1 (tcgen05 TF32) not built yet");
}
