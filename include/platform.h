// platform.h -- what the reference drivers expect from "platform.h" (ryg_rans platform.h:7-61):
// timer(), __rdtsc(), ALIGNSPEC and PRIu64.  Linux/x86-64 + GCC/Clang/NVCC host only; the
// GPU harness itself times with CUDA events (bench.py), this exists so main*.cpp build unchanged.
#ifndef PLATFORM_H_INCLUDED
#define PLATFORM_H_INCLUDED

#if !defined(__linux__) || !defined(__GNUC__)
#error "this package targets Linux hosts with a GNU-compatible compiler"
#endif

#ifndef __STDC_FORMAT_MACROS
#define __STDC_FORMAT_MACROS
#endif
#include <assert.h>
#include <inttypes.h>
#include <time.h>
#include <x86intrin.h>

#define ALIGNSPEC(type, name, alignment) type name __attribute__((aligned(alignment)))

// seconds on the monotonic clock
static inline double timer()
{
    struct timespec now;
    const int rc = clock_gettime(CLOCK_MONOTONIC, &now);
    assert(rc == 0);
    (void)rc;
    return (double)now.tv_sec + (double)now.tv_nsec * 1.0e-9;
}

#endif  // PLATFORM_H_INCLUDED
