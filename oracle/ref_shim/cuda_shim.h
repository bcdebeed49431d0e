// TEST INFRASTRUCTURE ONLY -- not product code.
//
// Host shim that lets g++ compile the reference's UNMODIFIED on-device CUDA
// sources (/root/reference/lib/on-device/*.cu) as plain C++, so the reference's
// own kernels can be run on CPU as the parity oracle and the CPU baseline
// (BASELINE.md section 3).  Force-included with
// `g++ -x c++ -include cuda_shim.h`.  No reference source is copied: the
// Makefile compiles the files where they lie under /root/reference.
#pragma once
#include <cmath>      // must precede the reference's `#define gamma (5.0 / 3.0)`
#include <cstdio>
#include <cstdlib>
#include <stdint.h>
#include <sys/types.h>  // u_int32_t (initialize_od.cu, kernels_od_intvar.cu)

#define __global__
#define __device__
#define __host__

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

// One emulated CUDA thread per host thread; defined in ref_harness.cpp.
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

static inline void __syncthreads() {}

// CUDA resolves pow(float,int) etc. through the C++ <cmath> overload set
// (float/int arguments promote to double), which is what std:: does too.
using std::cos;
using std::log;
using std::pow;
using std::sqrt;
