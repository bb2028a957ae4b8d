// One instantiation of the fused FP32 kernel, for quick SASS inspection while tuning:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I include -cubin -o /tmp/one.cubin tools/micro/one_kernel.cu
//   cuobjdump -sass /tmp/one.cubin
#define SART_NO_LAUNCHERS
#include "../../solaraxionraytracing_b200/csrc/kernels_f32.cu"
#ifndef ONE_ARGS
#define ONE_ARGS false, true, false, true   // cone optic, plain run, inverse-CDF sampler, margins
#endif
namespace sart { namespace fast {
void* one_kernel_address() { return reinterpret_cast<void*>(&k_trace_mc_f32<ONE_ARGS>); }
} }
