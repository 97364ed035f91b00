// ref_prelude.h -- shared by the oracle/_ref harness translation units.  TEST INFRASTRUCTURE ONLY.
//
// Each harness TU textually includes one UNMODIFIED reference driver
// (/root/reference/main*.cpp, located through -I at build time, never copied
// into this repo) inside its own namespace, with `main` renamed, so that the
// reference's SymbolStats / RansEnc* / RansDec* / RansWord* / Rans64* code is
// what actually runs.  System headers must be included first, outside the
// namespace; their include guards then turn the driver's own #includes into
// no-ops.
#pragma once
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <assert.h>
#include <time.h>
#ifndef __STDC_FORMAT_MACROS
#define __STDC_FORMAT_MACROS
#endif
#include <inttypes.h>
#include <x86intrin.h>
#include <smmintrin.h>
#include <new>
#include <thread>
#include <vector>
#include <algorithm>

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

static inline double ref_now()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}
