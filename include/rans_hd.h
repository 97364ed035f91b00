// rans_hd.h -- host/device decoration shared by the source-level rANS headers.
#ifndef RANS_HD_H
#define RANS_HD_H
#if defined(__CUDACC__)
#define RANS_HD __host__ __device__ __forceinline__
#define RANS_HDM __host__ __device__ __forceinline__   /* static member functions */
#else
#define RANS_HD static inline
#define RANS_HDM inline
#endif
#endif
