// Micro-benchmark for the contact-map inner loop (cmap_kernels.cu): what does one (row, column) pair cost on sm_100a for the
// candidate instruction sequences?  Every variant evaluates the same 32-row x 128-column unit (rows broadcast from shared memory,
// four columns per lane in registers) `iters` times per warp, 8 CTAs of 4 warps per SM, and prints pair evaluations per clock
// and SM (128 = one pair per lane and cycle would be the issue limit of a 1-instruction body).
//   0  scalar FSUB/FMUL/FADD (unfused) + FSETP + predicated LOP          (the shipped body, 10 instructions per pair)
//   1  scalar FP + integer compare: IADD (d - thr2 as bit patterns) + SHF.L funnel (rows walked 31..0)
//   2  packed f32x2 subtract / multiply, adds as fma(x, 1.0, y) (ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2, which
//      would not be bit-exact), FSETP + LOP
//   3  packed f32x2 + integer compare
//   4  packed subtract / multiply, scalar adds, FSETP + LOP
//   10..15  raw throughput of single instructions: FADD, FMUL x*x, FFMA, FADD2, FMUL2, FSETP+LOP pair
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/cmap_body_bench tools/cmap_body_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 lds_row(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64 pack2v(float lo, float hi)                     // volatile: packed once, not re-materialised per use
{
    u64 r;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int V>
__global__ void __launch_bounds__(128, 8) body_kernel(const float4 *__restrict__ pts, int iters, float thr2, float one, uint32_t *__restrict__ out)
{
    __shared__ float4 rows[32];
    __shared__ __align__(16) float2 rows2[32 * 4];                                        // (x,x) (y,y) (z,z) per row for the packed variants
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const float4 v = pts[threadIdx.x];
        rows[threadIdx.x] = v;
        rows2[4 * threadIdx.x] = make_float2(v.x, v.x);
        rows2[4 * threadIdx.x + 1] = make_float2(v.y, v.y);
        rows2[4 * threadIdx.x + 2] = make_float2(v.z, v.z);
    }
    __syncthreads();
    const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows), rows2_s = (uint32_t)__cvta_generic_to_shared(rows2);
    uint32_t acc = 0;
    const uint32_t thrb = __float_as_uint(thr2);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        float cx[4], cy[4], cz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = pts[32 + ((it * 4 + k) & 31) * 32 + lane];
            cx[k] = v.x; cy[k] = v.y; cz[k] = v.z;
        }
        uint32_t tw[4] = {0u, 0u, 0u, 0u};
        if (V == 0) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float4 a = lds_row(rows_s + 16 * r);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (sqdist3(a.x, a.y, a.z, cx[k], cy[k], cz[k]) < thr2) tw[k] |= 1u << r;
            }
        } else if (V == 1) {
#pragma unroll
            for (int r = 31; r >= 0; --r) {
                const float4 a = lds_row(rows_s + 16 * r);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t diff = __float_as_uint(sqdist3(a.x, a.y, a.z, cx[k], cy[k], cz[k])) - thrb;
                    tw[k] = __funnelshift_l(diff, tw[k], 1);
                }
            }
        } else {
            u64 px[2], py[2], pz[2];
            const u64 one2 = pack2(one, one);                                  // a run-time 1.0: ptxas must not simplify fma(x, 1, y) and contract
            // + a run-time 64-bit zero puts each column pair into a register pair of its own: ptxas otherwise rebuilds the pair
            // from the LDG destination registers with two MOVs in front of every use
            const u64 zero64 = (u64)(__float_as_uint(one) - 0x3f800000u);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                px[h] = pack2(cx[2 * h], cx[2 * h + 1]) + zero64; py[h] = pack2(cy[2 * h], cy[2 * h + 1]) + zero64;
                pz[h] = pack2(cz[2 * h], cz[2 * h + 1]) + zero64;
            }
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
                const int r = V == 3 ? 31 - rr : rr;
                u64 ax, ay, az;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(ax), "=l"(ay) : "r"(rows2_s + 32 * r));
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(az) : "r"(rows2_s + 32 * r + 16));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const u64 dx = sub2(ax, px[h]), dy = sub2(ay, py[h]), dz = sub2(az, pz[h]);
                    const u64 mx = mul2(dx, dx), my = mul2(dy, dy), mz = mul2(dz, dz);
                    float d0, d1;
                    if (V == 4) {
                        float x0, x1, y0, y1, z0, z1;
                        unpack2(mx, x0, x1); unpack2(my, y0, y1); unpack2(mz, z0, z1);
                        d0 = __fadd_rn(__fadd_rn(x0, y0), z0);
                        d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
                    } else {
                        const u64 d = fma2(fma2(mx, one2, my), one2, mz);
                        unpack2(d, d0, d1);
                    }
                    if (V == 3) {
                        tw[2 * h] = __funnelshift_l(__float_as_uint(d0) - thrb, tw[2 * h], 1);
                        tw[2 * h + 1] = __funnelshift_l(__float_as_uint(d1) - thrb, tw[2 * h + 1], 1);
                    } else {
                        if (d0 < thr2) tw[2 * h] |= 1u << r;
                        if (d1 < thr2) tw[2 * h + 1] |= 1u << r;
                    }
                }
            }
        }
        acc ^= tw[0] ^ (tw[1] * 3u) ^ (tw[2] * 5u) ^ (tw[3] * 7u);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// raw instruction throughput: 16 independent chains per thread
template <int OP>
__global__ void __launch_bounds__(128, 8) op_kernel(int iters, float seed, float *__restrict__ out)
{
    float x[16];
    u64 y[8];
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = seed + i + threadIdx.x; w[i] = i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) y[i] = pack2(x[2 * i], x[2 * i + 1]);
    const float c = seed * 0.5f;
    const u64 c2 = pack2(c, c);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (OP == 10) x[i] = __fadd_rn(x[i], c);
                if (OP == 11) x[i] = __fmul_rn(x[i], x[i]);
                if (OP == 12) x[i] = __fmaf_rn(x[i], c, x[(i + 1) & 15]);
                if (OP == 13 && i < 8) y[i] = add2(y[i], c2);
                if (OP == 14 && i < 8) y[i] = mul2(y[i], y[i]);
                if (OP == 15) { if (x[i] < c) w[i] |= 1u << rep; x[i] = __fadd_rn(x[i], c); }
                if (OP == 16) { w[i] = __funnelshift_l(__float_as_uint(x[i]) - __float_as_uint(c), w[i], 1); x[i] = __fadd_rn(x[i], c); }
                if (OP == 17) x[i] = __fmul_rn(x[i], c);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i] + (float)w[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; unpack2(y[i], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    setvbuf(stdout, nullptr, _IONBF, 0);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, grid = sms * 8;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float4 *pts; uint32_t *out; float *fout;
    cudaMalloc(&pts, sizeof(float4) * (32 + 32 * 32));
    cudaMalloc(&out, sizeof(uint32_t) * grid * 128);
    cudaMalloc(&fout, sizeof(float) * grid * 128);
    float4 h[32 + 32 * 32];
    srand(1);
    for (auto &v : h) v = make_float4(rand() % 2000 * 0.01f, rand() % 2000 * 0.01f, rand() % 2000 * 0.01f, 0.f);
    cudaMemcpy(pts, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 400;
    uint32_t ref[4] = {0, 0, 0, 0};
    auto run_body = [&](int v, auto kernel) {
        float ms = time_ms([&] { kernel<<<grid, 128>>>(pts, iters, 150.0f, 1.0f, out); });
        static uint32_t hout[148 * 8 * 128 * 2];
        cudaMemcpy(hout, out, sizeof(uint32_t) * grid * 128, cudaMemcpyDeviceToHost);
        uint32_t chk[4] = {0, 0, 0, 0};
        for (int i = 0; i < grid * 128; ++i) { chk[i & 3] ^= hout[i] * (uint32_t)(i + 1); chk[(i >> 2) & 3] += hout[i]; }
        if (v == 0) for (int i = 0; i < 4; ++i) ref[i] = chk[i];
        const double pairs = (double)grid * 4 * iters * 32 * 128;
        printf("body %d: %.3f ms, %.2f Tpairs/s, checksum %s (%08x)  [%s]\n", v, ms, pairs / ms / 1e9,
               (chk[0] == ref[0] && chk[1] == ref[1] && chk[2] == ref[2] && chk[3] == ref[3]) ? "same" : "DIFFERENT", chk[0], cudaGetErrorString(cudaGetLastError()));
    };
    run_body(0, body_kernel<0>);
    run_body(1, body_kernel<1>);
    run_body(2, body_kernel<2>);
    run_body(3, body_kernel<3>);
    run_body(4, body_kernel<4>);
    auto run_op = [&](int op, auto kernel, double per_iter) {
        float ms = time_ms([&] { kernel<<<grid, 128>>>(2000, 1.5f, fout); });
        const double n = (double)grid * 128 * 2000 * per_iter;
        printf("op %d: %.3f ms, %.2f T lane-instr/s (nominal clock %.3f GHz -> %.1f per clk and SM at that clock)\n", op, ms, n / ms / 1e9,
               khz / 1e6, n / ms / 1e3 / (khz * 1e3) / sms * 1e0);
    };
    run_op(10, op_kernel<10>, 128);
    run_op(11, op_kernel<11>, 128);
    run_op(17, op_kernel<17>, 128);
    run_op(12, op_kernel<12>, 128);
    run_op(13, op_kernel<13>, 64);
    run_op(14, op_kernel<14>, 64);
    run_op(15, op_kernel<15>, 128 * 3);
    run_op(16, op_kernel<16>, 128 * 3);
    return 0;
}
