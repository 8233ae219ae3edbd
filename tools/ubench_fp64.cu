// Micro-benchmarks behind the chain kernel's latency model (run on the B200 box):
//   dependent DFMA latency, rcp.approx.ftz.f64 latency + accuracy after k Newton/Halley steps,
//   shared-memory publish -> __syncwarp -> broadcast read round trip.
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__global__ void k_lat(double* out, long long* cyc, double x0) {
    __shared__ double sm[64];
    const int lane = threadIdx.x;
    double x = x0 + lane * 1e-3, y = 1.0000001;
    long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
    long long c1 = clock64();
    if (lane == 0) cyc[0] = c1 - c0;                       // 4096 dependent DFMA
    double r = x0 + 3.0 + lane;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q + 1.5; }
    c1 = clock64();
    if (lane == 0) cyc[1] = c1 - c0;                       // 1024 x (rcp.approx + DADD)
    double v = x;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { sm[lane] = v; __syncwarp(); v = sm[(lane + 1) & 31] + 1.0; __syncwarp(); }
    c1 = clock64();
    if (lane == 0) cyc[2] = c1 - c0;                       // 1024 x (STS, syncwarp, LDS, DADD, syncwarp)
    double d = x0 + 2.0;
    c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) { d = 1.0 / (d + 1.25); }
    c1 = clock64();
    if (lane == 0) cyc[3] = c1 - c0;                       // 1024 x (IEEE division + DADD)
    out[lane] = x + r + v + d;
}

__global__ void k_acc(double* err) {
    // accuracy of rcp.approx.ftz.f64 raw, +1 Newton, +2 Newton, +1 Halley(cubic), +Newton+Halley
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    double e[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4096; ++i) {
        unsigned long long h = (unsigned long long)(tid * 4096 + i) * 0x9E3779B97F4A7C15ull;
        double x = 1.0 + (double)(h >> 11) * (1.0 / 9007199254740992.0);   // [1,2)
        x = ldexp(x, (int)(h % 41) - 20);
        double r0; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
        const double t = 1.0 / x;
        double ee = fma(-x, r0, 1.0); double r1 = fma(r0, ee, r0);
        ee = fma(-x, r1, 1.0); double r2 = fma(r1, ee, r1);
        ee = fma(-x, r0, 1.0); double e2 = fma(ee, ee, ee); double rh = fma(r0, e2, r0);
        ee = fma(-x, r1, 1.0); e2 = fma(ee, ee, ee); double rnh = fma(r1, e2, r1);
        double c[5] = {r0, r1, r2, rh, rnh};
        for (int k = 0; k < 5; ++k) e[k] = fmax(e[k], fabs(c[k] - t) / t);
    }
    for (int k = 0; k < 5; ++k) atomicMax((unsigned long long*)&err[k], (unsigned long long)__double_as_longlong(e[k]));
}

int main() {
    double *out, *err; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8 * 8); cudaMalloc(&err, 5 * 8);
    cudaMemset(err, 0, 40);
    k_lat<<<1, 32>>>(out, cyc, 1.0);
    k_acc<<<64, 128>>>(err);
    long long h[4]; double he[5];
    cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost); cudaMemcpy(he, err, 40, cudaMemcpyDeviceToHost);
    printf("dependent DFMA latency        : %.1f cycles\n", h[0] / 4096.0);
    printf("rcp.approx.ftz.f64 + DADD     : %.1f cycles\n", h[1] / 1024.0);
    printf("STS/syncwarp/LDS/DADD/syncwarp: %.1f cycles\n", h[2] / 1024.0);
    printf("IEEE 1/x + DADD               : %.1f cycles\n", h[3] / 1024.0);
    printf("rcp.approx rel err: raw %.3e | +1 Newton %.3e | +2 Newton %.3e | +1 Halley %.3e | Newton+Halley %.3e\n", he[0], he[1], he[2], he[3], he[4]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
