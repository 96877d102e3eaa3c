// Instruction-throughput microbenchmark for sm_100a (development tool): cycles per warp
// instruction per SMSP for the ops the Lloyd kernel is made of, alone and mixed.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHAINS 8
#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float *out, uint32_t seed, long long *cycles) {
	float a[CHAINS], b[CHAINS];
	unsigned long long p[CHAINS], q[CHAINS];
	uint32_t u[CHAINS];
	float s = __int_as_float(0x3f800001 + seed), t = __int_as_float(0x3f000001 + seed);
	unsigned long long s2, t2;
	asm volatile("mov.b64 %0, {%1,%2};" : "=l"(s2) : "f"(s), "f"(t));
	asm volatile("mov.b64 %0, {%1,%2};" : "=l"(t2) : "f"(t), "f"(s));
	uint32_t m = 0xfffffff0u + seed;
#pragma unroll
	for (int i = 0; i < CHAINS; i++) {
		a[i] = threadIdx.x * 1e-3f + i; b[i] = i * 0.5f + seed;
		asm volatile("mov.b64 %0, {%1,%2};" : "=l"(p[i]) : "f"(a[i]), "f"(b[i]));
		q[i] = p[i] ^ 0x1000; u[i] = threadIdx.x + i;
	}
	__syncthreads();
	long long c0 = clock64();
	for (int it = 0; it < ITERS; it++) {
#pragma unroll
		for (int i = 0; i < CHAINS; i++) {
			if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(b[i]));          // FFMA 3 regs
			if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]));       // FFMA2 3 pairs
			if (MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));        // LOP3
			if (MODE == 3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s));            // FMNMX3
			if (MODE == 4) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));                        // FMNMX
			if (MODE == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(s2));                     // FADD2
			if (MODE == 6) {  // FFMA2 + LOP3 1:1
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));
			}
			if (MODE == 7) {  // the K-loop mix per centre pair: 3 FFMA2 + 1 FADD2 + 2 LOP3 + 1 FMNMX3
				asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(s2));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(t2), "l"(q[i]));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[(i + 1) % CHAINS]));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[(i + 1) % CHAINS]) : "r"(m), "n"(6));
				asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s));
			}
			if (MODE == 8) {  // scalar equivalent: 6 FFMA + 2 FADD + 2 LOP3 + 1 FMNMX3
				asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
				asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(b[i]) : "f"(s));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(t));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(b[i]) : "f"(s), "f"(t));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(t), "f"(s));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(b[i]) : "f"(t), "f"(s));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(b[(i + 1) % CHAINS]));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(b[i]) : "f"(s), "f"(a[(i + 1) % CHAINS]));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[(i + 1) % CHAINS]) : "r"(m), "n"(6));
				asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s));
			}
			if (MODE == 9) {  // FFMA scalar + LOP3 1:1
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(b[i]));
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));
			}
			if (MODE == 10) {  // FFMA2 + FFMA scalar 1:1
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(s), "f"(b[i]));
			}
			if (MODE == 11) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(a[i]) : "f"(s));                   // FFMA 2 distinct regs
			if (MODE == 12) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(a[i]) : "f"(s));            // FFMA imm
			if (MODE == 13) {  // FMNMX3 + FFMA2 1:1
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[i]));
				asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s));
			}
			if (MODE == 14) asm volatile("max.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(m));                          // IMNMX
			if (MODE == 15) {  // LOP3 + FMNMX3 (both ALU?)
				asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(u[i]) : "r"(m), "n"(5));
				asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(s));
			}

			if (MODE == 20) asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(u[i]) : "r"(m));                    // IMAD imm
			if (MODE == 21) asm volatile("add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(m));                           // IADD
			if (MODE == 22) asm volatile("shl.b32 %0, %0, 4; add.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(m));          // SHL+ADD (LEA?)
			if (MODE == 23) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(a[i]) : "f"(b[i])); // FSETP+FSEL
			if (MODE == 24) asm volatile("min.s32 %0, %0, %1; min.s32 %0, %0, %2;" : "+r"(u[i]) : "r"(m), "r"(u[(i+1)%CHAINS]));   // 3-input int min?
			if (MODE == 25) {  // new K-loop, packed: per pair 3 FFMA2 + 2 IMAD + 2 IMNMX
				asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(s2), "l"(q[i]), "l"(t2));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(t2), "l"(q[i]));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[(i + 1) % CHAINS]));
				uint32_t lo, hi; asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p[i]));
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(lo) : "r"(m));
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(hi) : "r"(m));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(lo));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(hi));
			}
			if (MODE == 26) {  // new K-loop, scalar: per pair 6 FFMA + 2 IMAD + 2 IMNMX
				float d0, d1;
				asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d0) : "f"(s), "f"(a[i]), "f"(t));
				asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d1) : "f"(s), "f"(b[i]), "f"(t));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d0) : "f"(t), "f"(a[i]));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d1) : "f"(t), "f"(b[i]));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d0) : "f"(s), "f"(b[(i + 1) % CHAINS]));
				asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d1) : "f"(s), "f"(a[(i + 1) % CHAINS]));
				uint32_t lo = __float_as_uint(d0), hi = __float_as_uint(d1);
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(lo) : "r"(m));
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(hi) : "r"(m));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(lo));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(hi));
			}
			if (MODE == 27) {  // TIE int tracking per pair: 3 FFMA2 + 2 IMAD + 6 IMNMX
				asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(s2), "l"(q[i]), "l"(t2));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(t2), "l"(q[i]));
				asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(s2), "l"(q[(i + 1) % CHAINS]));
				uint32_t lo, hi, mn, mx, tt; asm volatile("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p[i]));
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(lo) : "r"(m));
				asm volatile("mad.lo.s32 %0, %0, 16, %1;" : "+r"(hi) : "r"(m));
				asm volatile("min.s32 %0, %1, %2;" : "=r"(mn) : "r"(lo), "r"(hi));
				asm volatile("max.s32 %0, %1, %2;" : "=r"(mx) : "r"(lo), "r"(hi));
				asm volatile("max.s32 %0, %1, %2;" : "=r"(tt) : "r"(u[i]), "r"(mn));
				uint32_t sec = __float_as_uint(a[i]);
				asm volatile("min.s32 %0, %0, %1;" : "+r"(sec) : "r"(tt));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(sec) : "r"(mx));
				asm volatile("min.s32 %0, %0, %1;" : "+r"(u[i]) : "r"(mn));
				a[i] = __uint_as_float(sec);
			}
			if (MODE == 28) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(u[i]) : "r"(m));                  // PRMT
			if (MODE == 29) asm volatile("shf.l.wrap.b32 %0, %0, %1, 4;" : "+r"(u[i]) : "r"(m));                 // SHF
			if (MODE == 30) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(m), "r"(u[(i+1)%CHAINS])); // IMAD r,r,r
			if (MODE == 31) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(a[i]) : "r"(u[i]));                      // I2F
			if (MODE == 32) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));                        // FADD
			if (MODE == 33) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));                        // FMUL
		}
	}
	long long c1 = clock64();
	float r = 0;
#pragma unroll
	for (int i = 0; i < CHAINS; i++) r += a[i] + b[i] + (float)(p[i] & 0xff) + (float)u[i] + (float)(q[i] & 1);
	out[blockIdx.x * blockDim.x + threadIdx.x] = r;
	if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
}

template <int MODE> void run(const char *name, int per_iter, int threads) {
	float *out; long long *cyc;
	cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
	bench<MODE><<<148, threads>>>(out, 0, cyc);
	bench<MODE><<<148, threads>>>(out, 0, cyc);
	cudaDeviceSynchronize();
	long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
	double warps_per_smsp = threads / 32 / 4.0;
	double inst = (double)ITERS * CHAINS * per_iter * warps_per_smsp;
	printf("%-44s threads=%4d cycles=%9.0f  cyc/warp-inst/SMSP=%.3f\n", name, threads, avg, avg / inst);
	cudaFree(out); cudaFree(cyc);
}

int main() {
	for (int threads : {512}) {
		run<0>("FFMA r,r,r (3 distinct)", 1, threads);
		run<11>("FFMA (2 distinct regs)", 1, threads);
		run<12>("FFMA imm", 1, threads);
		run<1>("FFMA2", 1, threads);
		run<5>("FADD2", 1, threads);
		run<2>("LOP3", 1, threads);
		run<3>("FMNMX3", 1, threads);
		run<4>("FMNMX", 1, threads);
		run<14>("IMNMX", 1, threads);
		run<6>("FFMA2+LOP3 (2 inst)", 2, threads);
		run<9>("FFMA+LOP3 (2 inst)", 2, threads);
		run<10>("FFMA2+FFMA (2 inst)", 2, threads);
		run<13>("FFMA2+FMNMX3 (2 inst)", 2, threads);
		run<15>("LOP3+FMNMX3 (2 inst)", 2, threads);
		run<7>("K-loop mix packed (7 inst)", 7, threads);
		run<8>("K-loop mix scalar (11 inst)", 11, threads);
		run<20>("IMAD imm", 1, threads);
		run<30>("IMAD r,r,r", 1, threads);
		run<21>("IADD", 1, threads);
		run<22>("SHL+ADD (2 ptx ops)", 1, threads);
		run<29>("SHF", 1, threads);
		run<28>("PRMT", 1, threads);
		run<23>("FSETP+FSEL (as 1)", 1, threads);
		run<24>("min,min int (as 1)", 1, threads);
		run<31>("I2F", 1, threads);
		run<32>("FADD", 1, threads);
		run<33>("FMUL", 1, threads);
		run<25>("new K-loop packed (7 inst/pair)", 7, threads);
		run<26>("new K-loop scalar (10 inst/pair)", 10, threads);
		run<27>("TIE int packed (11 inst/pair)", 11, threads);
	}
	return 0;
}
