// Pair / triple instruction-mix microbenchmark for sm_100a (development tool): which of the Lloyd kernel's
// ops overlap when interleaved?  cycles per GROUP per SMSP, 4 warps per SMSP, 8 independent chains.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CH 8
#define ITERS 2048

#define F2(i)   asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]))
#define F1(i)   asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(b[i]))
#define F1b(i)  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(b[i]) : "f"(t), "f"(a[i]))
#define LEA_(i) asm volatile("{.reg .b32 tt; shl.b32 tt, %0, 4; add.s32 %0, tt, %1;}" : "+r"(u[i]) : "r"(m))
#define LEAv(i) asm volatile("{.reg .b32 tt; shl.b32 tt, %1, 4; add.s32 %0, tt, %2;}" : "=r"(v[i]) : "r"(u[i]), "r"(m))
#define IMN(i)  asm volatile("min.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(v[i]))
#define IMN3(i) asm volatile("min.u32 %0, %0, %1; min.u32 %0, %0, %2;" : "+r"(u[i]) : "r"(v[i]), "r"(m))
#define IAD(i)  asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(m))
#define FAD(i)  asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s))
#define FMN(i)  asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]))
#define LOP(i)  asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5))
#define PRM(i)  asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[i]) : "r"(m))
#define LDSX(i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b[i]) : "r"(saddr + 4 * i))

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float *out, uint32_t seed, long long *cycles) {
	__shared__ float sm[64];
	float a[CH], b[CH];
	unsigned long long p[CH], q[CH];
	uint32_t u[CH], v[CH];
	const uint32_t tz = threadIdx.x >> 10;  // 0, but not to the compiler: keeps the operands in vector registers
	float s = __int_as_float(0x3f800001 + seed + tz), t = __int_as_float(0x3f000001 + seed + tz);
	unsigned long long s2;
	asm volatile("mov.b64 %0, {%1,%2};" : "=l"(s2) : "f"(s), "f"(t));
	uint32_t m = 0xfffffff0u + seed + tz;
	if (threadIdx.x < 64) sm[threadIdx.x] = threadIdx.x;
	uint32_t saddr = (uint32_t)__cvta_generic_to_shared(sm);
#pragma unroll
	for (int i = 0; i < CH; i++) {
		a[i] = threadIdx.x * 1e-3f + i; b[i] = i * 0.5f + seed;
		asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
		q[i] = p[i] ^ 0x1000; u[i] = threadIdx.x + i; v[i] = threadIdx.x * 3 + i;
	}
	__syncthreads();
	long long c0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < CH; i++) {
			if (MODE == 0) { F2(i); LEA_(i); }
			if (MODE == 1) { F2(i); IMN(i); }
			if (MODE == 2) { F2(i); IMN3(i); }
			if (MODE == 3) { F2(i); IAD(i); }
			if (MODE == 4) { F1(i); IMN(i); }
			if (MODE == 5) { F1(i); IAD(i); }
			if (MODE == 6) { F1(i); LEA_(i); }
			if (MODE == 7) { LEA_(i); IMN(i); }
			if (MODE == 8) { LEAv(i); IMN3(i); }
			if (MODE == 9) { IMN(i); IAD(i); }
			if (MODE == 10) { F2(i); FAD(i); }
			if (MODE == 11) { F2(i); PRM(i); }
			if (MODE == 12) { F2(i); F2(i); F2(i); LEAv(i); LEAv(i); IMN3(i); }          // grouped (v written twice: fine)
			if (MODE == 13) { F2(i); LEAv(i); F2(i); LEAv(i); F2(i); IMN3(i); }          // interleaved
			if (MODE == 14) { F2(i); IMN3(i); F2(i); LEAv(i); F2(i); LEAv(i); }
			if (MODE == 15) { F2(i); F2(i); F2(i); LEAv(i); LEAv(i); IMN(i); IMN(i); }   // 2-input mins
			if (MODE == 16) { F1(i); F1b(i); F1(i); F1b(i); F1(i); F1b(i); LEAv(i); LEAv(i); IMN(i); IMN(i); }  // scalar FMAs
			if (MODE == 17) { F1(i); LEAv(i); F1b(i); IMN(i); F1(i); LEAv(i); F1b(i); IMN(i); F1(i); F1b(i); }
			if (MODE == 18) { F2(i); LDSX(i); }
			if (MODE == 19) { LEA_(i); LDSX(i); }
			if (MODE == 20) { F1(i); F1b(i); IMN(i); }
			if (MODE == 21) { F1(i); F1b(i); LEA_(i); }
			if (MODE == 22) { F1(i); F1b(i); F1(i); LEA_(i); IMN(i); }
			if (MODE == 23) { IMN3(i); }
			if (MODE == 24) { LEA_(i); }
			if (MODE == 25) { F2(i); F2(i); IMN3(i); }
			if (MODE == 26) { F2(i); F2(i); LEA_(i); }
			if (MODE == 27) { F2(i); F2(i); F2(i); IMN3(i); IMN3(i); IMN3(i); }
		}
	}
	long long c1 = clock64();
	float r = 0;
#pragma unroll
	for (int i = 0; i < CH; i++) r += a[i] + b[i] + (float)(p[i] & 0xff) + (float)u[i] + (float)v[i] + (float)(q[i] & 1);
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
	if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
}

template <int MODE> void run(const char *name, int threads) {
	float *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
	bench<MODE><<<148, threads>>>(out, 0, cyc);
	bench<MODE><<<148, threads>>>(out, 0, cyc);
	cudaDeviceSynchronize();
	long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
	double warps_per_smsp = threads / 32 / 4.0;
	printf("%-56s threads=%4d cyc/group/SMSP=%.3f\n", name, threads, avg / ((double)ITERS * CH * warps_per_smsp));
	cudaFree(out); cudaFree(cyc);
}

int main() {
	const int T = 512;
	run<24>("LEA", T); run<23>("VIMNMX3", T);
	run<0>("FFMA2 + LEA", T); run<1>("FFMA2 + IMNMX", T); run<2>("FFMA2 + VIMNMX3", T); run<3>("FFMA2 + IADD", T);
	run<4>("FFMA + IMNMX", T); run<5>("FFMA + IADD", T); run<6>("FFMA + LEA", T); run<7>("LEA + IMNMX", T);
	run<8>("LEA + VIMNMX3", T); run<9>("IMNMX + IADD", T); run<10>("FFMA2 + FADD", T); run<11>("FFMA2 + PRMT", T);
	run<25>("2 FFMA2 + VIMNMX3", T); run<26>("2 FFMA2 + LEA", T); run<27>("3 FFMA2 + 3 VIMNMX3", T);
	run<12>("3 FFMA2, 2 LEA, VIMNMX3 (grouped)        serial=12", T);
	run<13>("F2 LEA F2 LEA F2 VIMNMX3 (interleaved)   serial=12", T);
	run<14>("F2 VIMNMX3 F2 LEA F2 LEA                 serial=12", T);
	run<15>("3 FFMA2, 2 LEA, 2 IMNMX                  serial=12", T);
	run<16>("6 FFMA, 2 LEA, 2 IMNMX (grouped)         serial=12", T);
	run<17>("6 FFMA, 2 LEA, 2 IMNMX (interleaved)     serial=12", T);
	run<18>("FFMA2 + LDS", T); run<19>("LEA + LDS", T);
	run<20>("2 FFMA + IMNMX", T); run<21>("2 FFMA + LEA", T); run<22>("3 FFMA + LEA + IMNMX", T);
	return 0;
}
