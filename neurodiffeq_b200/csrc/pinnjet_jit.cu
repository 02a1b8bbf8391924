// pinnjet_jit.cu -- translation unit of the SPECIALISED forward kernel (neurodiffeq_b200/jit.py compiles it per problem with
// nvcc -cubin; it is not part of libpinnjet.so).  Two generated includes sit next to it in a temporary directory:
//   pinnjet_jit_scheme.inc    #define PJ_JIT_N1 / PJ_JIT_N2 / PJ_JIT_WL   (the jet-channel scheme of the problem)
//   pinnjet_jit_programs.inc  pj::pj_jit_program_train / _eval / _w       (the traced programs as straight-line CUDA)
// The kernel is k1tc3_body (pinnjet_k1tc3.cuh) with the interpreter calls replaced by those functions; same K1Args, same
// shared-memory plan, same workspace -- pj_forward_jit / pj_forward_train_jit (pinnjet_api.cu) launch it by handle.
#define PJ_JIT 1
#include "pinnjet_jit_scheme.inc"
#include "pinnjet_program.cuh"
#include "pinnjet_jit_programs.inc"
#include "pinnjet_k1tc3.cuh"

extern "C" __global__ void __launch_bounds__(pj::K1T_THREADS, 1) pj_k1_jit(const __grid_constant__ pj::K1Args A) {
    pj::k1tc3_body<PJ_JIT_N1, PJ_JIT_N2, PJ_JIT_WL, true>(A);
}
