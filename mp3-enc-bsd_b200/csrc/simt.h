// simt.h — one source, two compilations.
//
// The kernels in this directory are written in a phase-structured SIMT style: per-thread code lives
// inside FOR_THREADS(ctx) { ... } END_THREADS blocks, cross-thread communication happens only
// through shared memory + ctx.sync() or through the explicit warp collectives of the context.
//
//  * nvcc (device): PerThread<T> is a plain register, FOR_THREADS runs the body once for the calling
//    thread, collectives map to redux.sync / shfl.sync / bar.sync.  Zero overhead.
//  * g++ -DMP3GPU_HOST_EMUL (tests only, tests/emul): PerThread<T> is an array over the threads of
//    the group, FOR_THREADS is a loop, collectives are loops.  This lets `pytest -m "not gpu"` run
//    the *identical kernel logic* against the oracle on a machine without a GPU.  It is a test
//    harness, never a product fallback: libmp3gpu.so contains no host-emulated path.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MP3GPU_HOST_EMUL)
#define SIMT_DEV 1
#define SIMT_FN __device__ __forceinline__
#define SIMT_HD __host__ __device__ __forceinline__
#define SIMT_NOINLINE static __device__ __noinline__
#else
#define SIMT_DEV 0
#define SIMT_FN inline
#define SIMT_HD inline
#define SIMT_NOINLINE static inline
#include <cmath>
#include <cstring>
#endif

namespace simt {

#if SIMT_DEV
// ------------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------------
template <class T, int N = 32>
struct PerThread {
    T v;
    SIMT_FN T &operator()() { return v; }
    SIMT_FN const T &operator()() const { return v; }
};

struct WarpCtx {  // one warp == one group
    int lane;
    static constexpr int kThreads = 32;
    SIMT_FN WarpCtx() : lane(threadIdx.x & 31) {}
    // lane index that the optimiser cannot rematerialise from %tid inside hot loops (it lives in a register instead)
    struct Pinned {};
    SIMT_FN explicit WarpCtx(Pinned) { const int l = threadIdx.x & 31; lane = __shfl_sync(0xffffffffu, l, l); }
    SIMT_FN void sync() const { __syncwarp(); }
    SIMT_FN int reduce_max(const PerThread<int> &x) const { return __reduce_max_sync(0xffffffffu, x.v); }
    SIMT_FN int reduce_add(const PerThread<int> &x) const { return __reduce_add_sync(0xffffffffu, x.v); }
    SIMT_FN unsigned ballot(const PerThread<int> &x) const { return __ballot_sync(0xffffffffu, x.v != 0); }
    // maximum of non-negative floats (their bit patterns order like unsigned integers): one redux instead of a shuffle tree
    SIMT_FN float reduce_max_nonneg(const PerThread<float> &x) const { return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(x.v))); }
    SIMT_FN double reduce_max(const PerThread<double> &x) const {
        double v = x.v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    // fixed butterfly order: deterministic, identical in the host emulation
    SIMT_FN double reduce_add(const PerThread<double> &x) const {
        double v = x.v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
        return v;
    }
    // dst(lane) = src(lane + 1), lane 31 gets `wrap`(lane 0 of the next row, supplied by caller)
    SIMT_FN void shift_down1(PerThread<int> &dst, const PerThread<int> &src, const PerThread<int> &next_row) const {
        int a = __shfl_down_sync(0xffffffffu, src.v, 1);
        int b = __shfl_sync(0xffffffffu, next_row.v, 0);
        dst.v = (lane == 31) ? b : a;
    }
    template <class T>
    SIMT_FN T broadcast(const PerThread<T> &x, int src_lane) const { return __shfl_sync(0xffffffffu, x.v, src_lane); }
    // dst(lane) = src(lane + d) (lanes past the end keep their own value); dst(lane) = src(idx(lane))
    SIMT_FN void shfl_down(PerThread<double> &dst, const PerThread<double> &src, int d) const { dst.v = __shfl_down_sync(0xffffffffu, src.v, d); }
    SIMT_FN void shfl_idx(PerThread<double> &dst, const PerThread<double> &src, const PerThread<int> &idx) const { dst.v = __shfl_sync(0xffffffffu, src.v, idx.v); }
};

// Reference to a per-warp shared-memory object whose address is opaque to the optimiser.  Under register pressure
// ptxas otherwise re-derives the shared-window address from %tid / %cgaid in every trip of every hot loop (S2R + 6
// integer instructions per access site, profiles/r01_e_capture.md); a value that went through a shuffle has to stay in a
// register.  All lanes of the warp must pass the same object.
template <class T>
SIMT_FN T &pin_smem(T &m)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(&m);
    a = __shfl_sync(0xffffffffu, a, 0);
    return *reinterpret_cast<T *>(__cvta_shared_to_generic(a));
}

struct BlockCtx {  // a whole CTA
    int tid, nthreads;
    SIMT_FN BlockCtx() : tid(threadIdx.x), nthreads(blockDim.x) {}
    SIMT_FN void sync() const { __syncthreads(); }
};

#define FOR_THREADS(ctx) { const int lane = (ctx).lane; (void)lane;
#define END_THREADS }
#define FOR_BLOCK_THREADS(ctx) { const int tid = (ctx).tid; (void)tid;
#define END_BLOCK_THREADS }

SIMT_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
SIMT_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
SIMT_FN double dsub(double a, double b) { return __dsub_rn(a, b); }
SIMT_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
SIMT_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
SIMT_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
SIMT_FN float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SIMT_FN float fmin_(float a, float b) { return fminf(a, b); }                 // nan-suppressing: fmin_(nan, b) == b
// 2^23 + floor(a) for 0 <= a < 2^23, as a float whose low mantissa bits ARE floor(a): one add (round down) on the FMA pipe
// where a float -> int conversion would go through the quarter-rate XU pipe
SIMT_FN float floor_magic(float a) { return __fadd_rd(a, 8388608.0f); }
SIMT_FN unsigned fbits(float a) { return __float_as_uint(a); }
SIMT_FN unsigned pack_lo16(unsigned lo, unsigned hi) { return __byte_perm(lo, hi, 0x5410); }   // (lo & 0xffff) | (hi << 16)
SIMT_FN int f2i_trunc(float a) { return __float2int_rz(a); }   // saturating; nan -> 0
SIMT_FN int popc(unsigned v) { return __popc(v); }
SIMT_FN int clz(unsigned v) { return __clz(v); }
SIMT_FN unsigned umax(unsigned a, unsigned b) { return max(a, b); }
SIMT_FN unsigned vminu2(unsigned a, unsigned b) { return __vminu2(a, b); }   // per-halfword unsigned minimum
SIMT_FN unsigned vmaxu2(unsigned a, unsigned b) { return __vmaxu2(a, b); }   // per-halfword unsigned maximum

#else
// ------------------------------------------------------------------------------------------------
// host emulation (tests only)
// ------------------------------------------------------------------------------------------------
extern thread_local int g_tid;  // defined in tests/emul

template <class T, int N = 32>
struct PerThread {
    T v[N];
    T &operator()() { return v[g_tid]; }
    const T &operator()() const { return v[g_tid]; }
};

struct WarpCtx {
    int lane;  // unused on host (loop variable shadows it)
    static constexpr int kThreads = 32;
    WarpCtx() : lane(0) {}
    void sync() const {}
    int reduce_max(const PerThread<int> &x) const { int m = x.v[0]; for (int i = 1; i < 32; i++) if (x.v[i] > m) m = x.v[i]; return m; }
    int reduce_add(const PerThread<int> &x) const { int s = 0; for (int i = 0; i < 32; i++) s += x.v[i]; return s; }
    unsigned ballot(const PerThread<int> &x) const { unsigned b = 0; for (int i = 0; i < 32; i++) if (x.v[i]) b |= 1u << i; return b; }
    float reduce_max_nonneg(const PerThread<float> &x) const { float m = x.v[0]; for (int i = 1; i < 32; i++) if (x.v[i] > m) m = x.v[i]; return m; }
    double reduce_max(const PerThread<double> &x) const {
        double t[32]; for (int i = 0; i < 32; i++) t[i] = x.v[i];
        for (int o = 16; o > 0; o >>= 1) { double u[32]; for (int i = 0; i < 32; i++) u[i] = std::fmax(t[i], t[i ^ o]); std::memcpy(t, u, sizeof(t)); }
        return t[0];
    }
    double reduce_add(const PerThread<double> &x) const {
        double t[32]; for (int i = 0; i < 32; i++) t[i] = x.v[i];
        for (int o = 16; o > 0; o >>= 1) { double u[32]; for (int i = 0; i < 32; i++) u[i] = t[i] + t[i ^ o]; std::memcpy(t, u, sizeof(t)); }
        return t[0];
    }
    void shift_down1(PerThread<int> &dst, const PerThread<int> &src, const PerThread<int> &next_row) const {
        for (int i = 0; i < 31; i++) dst.v[i] = src.v[i + 1];
        dst.v[31] = next_row.v[0];
    }
    template <class T>
    T broadcast(const PerThread<T> &x, int src_lane) const { return x.v[src_lane]; }
    void shfl_down(PerThread<double> &dst, const PerThread<double> &src, int d) const { double t[32]; for (int i = 0; i < 32; i++) t[i] = src.v[i + d < 32 ? i + d : i]; std::memcpy(dst.v, t, sizeof(t)); }
    void shfl_idx(PerThread<double> &dst, const PerThread<double> &src, const PerThread<int> &idx) const { double t[32]; for (int i = 0; i < 32; i++) t[i] = src.v[idx.v[i] & 31]; std::memcpy(dst.v, t, sizeof(t)); }
};

#define FOR_THREADS(ctx) for (int lane = 0; lane < 32; ++lane) { simt::g_tid = lane;
#define END_THREADS }

inline double dmul(double a, double b) { return a * b; }
inline double dadd(double a, double b) { return a + b; }
inline double dsub(double a, double b) { return a - b; }
inline float fmul(float a, float b) { return a * b; }
inline float fadd(float a, float b) { return a + b; }
inline float fsub(float a, float b) { return a - b; }
inline float ffma(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float fmin_(float a, float b) { return std::fmin(a, b); }
inline float floor_magic(float a) { return (float)(std::floor((double)a) + 8388608.0); }
inline unsigned fbits(float a) { unsigned u; std::memcpy(&u, &a, 4); return u; }
inline unsigned pack_lo16(unsigned lo, unsigned hi) { return (lo & 0xffffu) | (hi << 16); }
inline int f2i_trunc(float a) { return (a != a) ? 0 : (a >= 2147483648.0f) ? 2147483647 : (a <= -2147483648.0f) ? (-2147483647 - 1) : (int)a; }
inline int popc(unsigned v) { return __builtin_popcount(v); }
inline int clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned vminu2(unsigned a, unsigned b)
{
    const unsigned lo = (a & 0xffffu) < (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu);
    const unsigned hi = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
inline unsigned vmaxu2(unsigned a, unsigned b)
{
    const unsigned lo = (a & 0xffffu) > (b & 0xffffu) ? (a & 0xffffu) : (b & 0xffffu);
    const unsigned hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return lo | (hi << 16);
}
#endif

}  // namespace simt
