// Experiment: cycles per packet of the serial queue recurrence (network_sim.py:72-84) for a lone warp,
// in several exact formulations.  nvcc -O3 -fmad=false -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct P { double d_bw, dl, w_full, max_qd, inv_rate; };

#define LOOP_HEAD                                                                   \
    double q = 0.01, tu = 0.0, tt = 0.001;                                          \
    unsigned long long m = mask;                                                    \
    double acc = 0.0;                                                               \
    long long t0 = clock64();                                                       \
    for (int k = 0; k < n; ++k) {                                                   \
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
#define LOOP_TAIL                                                                   \
        acc += ll;                                                                  \
        tt = tt + p.inv_rate;                                                       \
    }                                                                               \
    long long t1 = clock64();                                                       \
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = acc + q + tu; }

// V0: the product's current formulation (w via integer mask, tail-drop threshold on w)
__global__ void v0(P p, int n, unsigned long long mask, int active, double *out)
{
    if ((int)threadIdx.x >= active) return;
    LOOP_HEAD
        const long long yb = __double_as_longlong(q - (tt - tu));
        const double w = __longlong_as_double(yb & ~(yb >> 63));
        const double c = p.d_bw + w;
        const bool full = w > p.w_full;
        const double ll = p.dl + w;
        q = rdrop ? q : (full ? w : c);
        tu = rdrop ? tu : tt;
    LOOP_TAIL
}
// V1: reference-shaped (compare the sum against max_qd, as before the threshold trick)
__global__ void v1(P p, int n, unsigned long long mask, int active, double *out)
{
    if ((int)threadIdx.x >= active) return;
    LOOP_HEAD
        const double y = q - (tt - tu);
        const double w = y > 0.0 ? y : 0.0;
        const double c = p.d_bw + w;
        const bool full = c > p.max_qd;
        const double ll = p.dl + w;
        q = rdrop ? q : (full ? w : c);
        tu = rdrop ? tu : tt;
    LOOP_TAIL
}
// V2: x precomputed off the chain is what the compiler should do anyway; here the selects are predicated moves
__global__ void v2(P p, int n, unsigned long long mask, int active, double *out)
{
    if ((int)threadIdx.x >= active) return;
    LOOP_HEAD
        const double x = tt - tu;
        const double y = q - x;
        const double cpos = p.d_bw + y;                 // candidate if y > 0 and not full
        const bool pos = y > 0.0;
        const bool fullp = y > p.w_full;                // full test on y (valid when pos)
        const double w = pos ? y : 0.0;
        const double ll = p.dl + w;
        double qn = cpos;
        if (fullp) qn = y;
        if (!pos) qn = (0.0 > p.w_full) ? 0.0 : p.d_bw;
        if (!rdrop) { q = qn; tu = tt; }
    LOOP_TAIL
}
// V3: fmax instead of the integer mask
__global__ void v3(P p, int n, unsigned long long mask, int active, double *out)
{
    if ((int)threadIdx.x >= active) return;
    LOOP_HEAD
        const double w = fmax(q - (tt - tu), 0.0);
        const double c = p.d_bw + w;
        const bool full = w > p.w_full;
        const double ll = p.dl + w;
        q = rdrop ? q : (full ? w : c);
        tu = rdrop ? tu : tt;
    LOOP_TAIL
}
// V4: only the bare chain (no ll / tt side work) -- the dependency floor
__global__ void v4(P p, int n, unsigned long long mask, int active, double *out)
{
    if ((int)threadIdx.x >= active) return;
    double q = 0.01, x = 0.0009;
    unsigned long long m = mask;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        const long long yb = __double_as_longlong(q - x);
        const double w = __longlong_as_double(yb & ~(yb >> 63));
        const double c = p.d_bw + w;
        const bool full = w > p.w_full;
        q = rdrop ? q : (full ? w : c);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = q; }
}
// V5: latency probes: dependent DADD chain, dependent DSETP+SEL chain, dependent IADD chain
__global__ void v5(int n, double *out)
{
    double a = 1.0, b = 1e-9;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) a = a + b;
    long long t1 = clock64();
    double c = 1.0;
    for (int k = 0; k < n; ++k) c = (c > 0.5) ? c * 1.0000001 : c;   // DMUL + DSETP + SEL
    long long t2 = clock64();
    unsigned u = 1;
    for (int k = 0; k < n; ++k) u = u * 3u + 1u;
    long long t3 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = (double)(t2 - t1) / n; out[2] = (double)(t3 - t2) / n; out[3] = a + c + u; }
}


// ---- the real loop's ingredients, added one at a time to V2 ------------------------------------------------
#define V2_BODY(TK)                                                                 \
        const double y = q - ((TK) - tu);                                           \
        const double cpos = p.d_bw + y;                                             \
        const bool pos = y > 0.0;                                                   \
        const bool fullp = y > p.w_full;                                            \
        const double w = pos ? y : 0.0;                                             \
        const double ll = p.dl + w;                                                 \
        double qn = fullp ? y : cpos;                                               \
        qn = pos ? qn : k0;                                                         \
        const bool full = pos ? fullp : full0;                                      \
        q = rdrop ? q : qn;                                                         \
        tu = rdrop ? tu : (TK);                                                     \
        const bool dropped = rdrop || full;                                         \
        const double lsigned = __longlong_as_double(__double_as_longlong(ll) | (dropped ? (long long)0x8000000000000000ull : 0ll));

// V6: V2 + record staged to shared memory (STS.128), send time by recurrence, counted loop
__global__ void v6(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[64];
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, tt = 0.001; unsigned long long m = mask;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        V2_BODY(tt)
        stage[k & 63] = make_double2(tt + ll, lsigned);
        tt = tt + p.inv_rate;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = q + tu + stage[3].x; }
}
// V7: V6 but the send time comes from shared memory (as in the product's counted loop)
__global__ void v7(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[64];
    __shared__ double tts[64];
    if ((int)threadIdx.x >= active) return;
    for (int j = 0; j < 64; j++) tts[j] = 0.001 + j * p.inv_rate;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0; unsigned long long m = mask;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        const double tk = tts[k & 63];
        V2_BODY(tk)
        stage[k & 63] = make_double2(tk + ll, lsigned);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = q + tu + stage[3].x; }
}
// V8: V6 with the data-dependent exit test of the original loop
__global__ void v8(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[64];
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, tt = 0.001; unsigned long long m = mask;
    const double end = 0.001 + (n + 0.5) * p.inv_rate;
    long long t0 = clock64();
    int k = 0;
    for (; k < 2 * n; ++k) {
        if (!(tt < end)) break;
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        V2_BODY(tt)
        stage[k & 63] = make_double2(tt + ll, lsigned);
        tt = tt + p.inv_rate;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / k; out[1] = q + tu + stage[3].x; }
}
// V9: V6 but records go straight to global memory (STG.128), as the per-lane path does
__global__ void v9(P p, int n, unsigned long long mask, int active, double *out, double2 *ring)
{
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, tt = 0.001; unsigned long long m = mask;
    double2 *r = ring + (size_t)threadIdx.x * 65536;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        V2_BODY(tt)
        r[k & 65535] = make_double2(tt + ll, lsigned);
        tt = tt + p.inv_rate;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = q + tu; }
}
// V10: chunked like the product: per 64 packets a counted pre-loop for the send times, then the main loop
__global__ void v10(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[64];
    __shared__ double tts[66];
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, t = 0.001;
    const double end = 1e30;
    long long t0 = clock64();
    for (int c = 0; c < n / 64; ++c) {
        unsigned long long m = mask + c;
        double tt = t; int cnt = 0;
#pragma unroll 8
        for (int j = 0; j < 64; ++j) { tts[j] = tt; cnt += (tt < end) ? 1 : 0; tt = tt + p.inv_rate; }
        tts[64] = tt;
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
            const double tk = tts[k];
            const bool rdrop = (m & 1ull) != 0ull; m >>= 1;
            V2_BODY(tk)
            stage[k] = make_double2(tk + ll, lsigned);
        }
        t = tts[cnt];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / (n / 64 * 64); out[1] = q + tu + stage[3].x; }
}

// V11: V6 with the record store software-pipelined by one packet (record of k-1 finished during k's chain)
__global__ void v11(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[66];
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, tt = 0.001; unsigned long long m = mask;
    double pw = 0.0, pt = 0.0; bool pd = false;
    long long t0 = clock64();
    for (int k = 0; k < n; ++k) {
        const bool rdrop = (m & 1ull) != 0ull; m = (m >> 1) | (m << 63);
        const double y = q - (tt - tu);
        // record of the previous packet (independent of this packet's chain)
        const double pll = p.dl + pw;
        const double pls = __longlong_as_double(__double_as_longlong(pll) | (pd ? (long long)0x8000000000000000ull : 0ll));
        stage[(k + 63) & 63] = make_double2(pt + pll, pls);
        const double cpos = p.d_bw + y;
        const bool pos = y > 0.0;
        const bool fullp = y > p.w_full;
        const double w = pos ? y : 0.0;
        double qn = fullp ? y : cpos;
        qn = pos ? qn : k0;
        const bool full = pos ? fullp : full0;
        q = rdrop ? q : qn;
        tu = rdrop ? tu : tt;
        pw = w; pt = tt; pd = rdrop || full;
        tt = tt + p.inv_rate;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / n; out[1] = q + tu + stage[3].x + pw + pt + pd; }
}
// V12: chunked: count-only pre-loop (send-time recurrence, no stores), then V11 as a counted loop
__global__ void v12(P p, int n, unsigned long long mask, int active, double *out)
{
    __shared__ double2 stage[66];
    if ((int)threadIdx.x >= active) return;
    const double k0 = (0.0 > p.w_full) ? 0.0 : p.d_bw; const bool full0 = 0.0 > p.w_full;
    double q = 0.01, tu = 0.0, t = 0.001;
    const double end = 1e30;
    long long t0 = clock64();
    for (int c = 0; c < n / 64; ++c) {
        unsigned long long m = mask + c;
        double tt = t; int cnt = 0;
#pragma unroll 16
        for (int j = 0; j < 64; ++j) { cnt += (tt < end) ? 1 : 0; tt = tt + p.inv_rate; }
        tt = t;
        double pw = 0.0, pt = 0.0; bool pd = false;
#pragma unroll 2
        for (int k = 0; k < cnt; ++k) {
            const bool rdrop = (m & 1ull) != 0ull; m >>= 1;
            const double y = q - (tt - tu);
            const double pll = p.dl + pw;
            const double pls = __longlong_as_double(__double_as_longlong(pll) | (pd ? (long long)0x8000000000000000ull : 0ll));
            stage[k] = make_double2(pt + pll, pls);          // slot k holds packet k-1 (slot 0: dummy)
            const double cpos = p.d_bw + y;
            const bool pos = y > 0.0;
            const bool fullp = y > p.w_full;
            const double w = pos ? y : 0.0;
            double qn = fullp ? y : cpos;
            qn = pos ? qn : k0;
            const bool full = pos ? fullp : full0;
            q = rdrop ? q : qn;
            tu = rdrop ? tu : tt;
            pw = w; pt = tt; pd = rdrop || full;
            tt = tt + p.inv_rate;
        }
        const double pll = p.dl + pw;
        stage[cnt] = make_double2(pt + pll, pll);
        t = tt;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (double)(t1 - t0) / (n / 64 * 64); out[1] = q + tu + stage[3].x; }
}

int main()
{
    P p{1.0 / 300.0, 0.1, 0.02, 0.0233333, 0.0011};
    double *out; cudaMallocManaged(&out, 64);
    const int n = 200000;
    const unsigned long long mask = 0x0101010100010001ull;
    for (int active : {1, 4, 32}) {
        v0<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V0 current        %.1f cyc/pkt\n", active, out[0]);
        v1<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V1 reference-form %.1f cyc/pkt\n", active, out[0]);
        v2<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V2 speculative    %.1f cyc/pkt\n", active, out[0]);
        v3<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V3 fmax           %.1f cyc/pkt\n", active, out[0]);
        v4<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V4 bare chain     %.1f cyc/pkt\n", active, out[0]);
    }
    double2 *ring; cudaMalloc(&ring, 32 * 65536 * sizeof(double2));
    for (int active : {1, 32}) {
        v6<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V6 V2+STS              %.1f cyc/pkt\n", active, out[0]);
        v7<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V7 V2+STS+LDS tt       %.1f cyc/pkt\n", active, out[0]);
        v8<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V8 V2+STS+exit test    %.1f cyc/pkt\n", active, out[0]);
        v9<<<1, 32>>>(p, n, mask, active, out, ring); cudaDeviceSynchronize(); printf("active=%2d V9 V2+STG global       %.1f cyc/pkt\n", active, out[0]);
        v10<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V10 chunked (product)  %.1f cyc/pkt\n", active, out[0]);
        v11<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V11 pipelined store    %.1f cyc/pkt\n", active, out[0]);
        v12<<<1, 32>>>(p, n, mask, active, out); cudaDeviceSynchronize(); printf("active=%2d V12 chunked pipelined  %.1f cyc/pkt\n", active, out[0]);
    }
    v5<<<1, 32>>>(n, out); cudaDeviceSynchronize();
    printf("latency probes: DADD %.1f  DMUL+DSETP+SEL %.1f  IMAD %.1f cycles\n", out[0], out[1], out[2]);
    return 0;
}
