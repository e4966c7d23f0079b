// microbench.cu — measured issue rates of the integer instructions the ME kernels are made of.
//
// MEASURED_PEAKS.json carries only HBM and bf16 numbers; the roofline denominator of the search
// kernel (SURVEY.md §8(d) "INT peak") is measured here: every test runs 8 independent dependent
// chains per thread of one SASS instruction, 1024 threads per SM x 148 SMs, and reports
// warp-instructions per clock per SM (from in-kernel clock64) and lane-ops/s (from CUDA events).
// Output: one JSON object on stdout (written to profiles/INT_PEAKS_r01.json by the GPU run).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CHECK(x)                                                                  \
    do {                                                                          \
        cudaError_t e = (x);                                                      \
        if (e != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));               \
            exit(1);                                                              \
        }                                                                         \
    } while (0)

constexpr int ITERS = 2048;
constexpr int CH = 8;

enum Op { VSAD, IADD3, IMAD, LEA, VIMNMX, VIMNMX3, LOP3, PRMT, SHF, MIX_SAD_IMAD, MIX_SAD_IMAD_MIN, LDS32, LDS128,
          CREDUX, MIX_ME, N_OPS };
const char *kNames[N_OPS] = {"vabsdiff4_acc", "iadd3", "imad", "lea", "vimnmx_u32", "vimnmx3_u32", "lop3", "prmt", "shf",
                             "mix_2sad_1imad", "mix_64sad_66imad_41min", "lds32", "lds128", "credux_min", "mix_me_171"};

template <int OP>
__global__ void __launch_bounds__(256) k(unsigned *out, long long *cycles, unsigned seed)
{
    __shared__ uint4 sm[1024];
    unsigned a[CH], b = seed + threadIdx.x, c = seed * 3 + 1;
    for (int i = 0; i < CH; i++) a[i] = seed + i * 77 + threadIdx.x;
    if (OP == LDS32 || OP == LDS128)
        for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_uint4(i & 1023, i & 1023, i & 1023, i & 1023);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == VSAD) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 15; add.u32 %0, t, %1; }" : "+r"(a[i]) : "r"(b));
            if (OP == VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b + i));
            if (OP == VIMNMX3)
                asm volatile("{ .reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b + i), "r"(c + i));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3210;" : "+r"(a[i]) : "r"(b));
            if (OP == SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a[i]) : "r"(b));
            if (OP == MIX_SAD_IMAD) {
                asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
                if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            }
            if (OP == LDS32) a[i] = ((unsigned *)sm)[(a[i] + threadIdx.x) & 1023];
            if (OP == LDS128) {
                uint4 v = sm[(a[i] + threadIdx.x) & 1023];
                a[i] = v.x ^ v.y ^ v.z ^ v.w;
            }
            if (OP == CREDUX) a[i] = __reduce_min_sync(0xFFFFFFFFu, a[i] + i);
        }
        if (OP == MIX_SAD_IMAD_MIN || OP == MIX_ME) {
            // the per-candidate instruction mix of the search kernel: 64 SAD, 25 adds, 41 packs, 41 mins
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < CH; i++)
                    asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b + r), "r"(c));
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < CH; i++)
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c + r));
#pragma unroll
            for (int r = 0; r < 5; r++)
#pragma unroll
                for (int i = 0; i < CH; i++)
                    if (OP == MIX_ME)
                        asm volatile("{ .reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2; }"
                                     : "+r"(a[i]) : "r"(b + i), "r"(c + r));
                    else
                        asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b + r));
        }
    }
    long long t1 = clock64();
    unsigned s = 0;
    for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(int sms, unsigned *d_out, long long *d_cyc, std::string &json)
{
    const int blocks = sms * 4, threads = 256;   // 4 CTAs x 8 warps = 32 warps per SM, 8 per SMSP
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; w++) k<OP><<<blocks, threads>>>(d_out, d_cyc, 1);
    CHECK(cudaEventRecord(e0));
    k<OP><<<blocks, threads>>>(d_out, d_cyc, 1);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(blocks);
    CHECK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double avg = 0;
    for (long long v : cyc) avg += (double)v;
    avg /= blocks;
    double per_iter = CH;                        // warp-instructions per thread-iteration
    if (OP == MIX_SAD_IMAD) per_iter = CH * 1.5;
    if (OP == VIMNMX3) per_iter = CH;            // counted as SASS instructions if fused; reported per PTX pair
    if (OP == MIX_SAD_IMAD_MIN) per_iter = CH * (8 + 8 + 5);
    if (OP == MIX_ME) per_iter = CH * (8 + 8 + 5);
    const double warp_instr_per_sm = per_iter * ITERS * 32.0;     // 32 warps per SM
    const double ipc_sm = warp_instr_per_sm / avg;
    const double lane_ops_s = per_iter * ITERS * (double)blocks * threads / (ms * 1e-3);
    char buf[512];
    snprintf(buf, sizeof buf,
             "  \"%s\": {\"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.1f, "
             "\"tera_lane_ops_per_s\": %.3f, \"ms\": %.4f, \"avg_cycles\": %.0f, \"implied_mhz\": %.0f},\n",
             kNames[OP], ipc_sm, ipc_sm * 32, lane_ops_s * 1e-12, ms, avg, avg / (ms * 1e-3) * 1e-6);
    json += buf;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main()
{
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    unsigned *d_out;
    long long *d_cyc;
    CHECK(cudaMalloc(&d_out, sizeof(unsigned) * sms * 4 * 256));
    CHECK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 4));
    std::string json = "{\n";
    char head[512];
    snprintf(head, sizeof head, "  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", prop.name, sms, prop.clockRate);
    json += head;
    run<VSAD>(sms, d_out, d_cyc, json);
    run<IADD3>(sms, d_out, d_cyc, json);
    run<IMAD>(sms, d_out, d_cyc, json);
    run<LEA>(sms, d_out, d_cyc, json);
    run<VIMNMX>(sms, d_out, d_cyc, json);
    run<VIMNMX3>(sms, d_out, d_cyc, json);
    run<LOP3>(sms, d_out, d_cyc, json);
    run<PRMT>(sms, d_out, d_cyc, json);
    run<SHF>(sms, d_out, d_cyc, json);
    run<MIX_SAD_IMAD>(sms, d_out, d_cyc, json);
    run<MIX_SAD_IMAD_MIN>(sms, d_out, d_cyc, json);
    run<MIX_ME>(sms, d_out, d_cyc, json);
    run<LDS32>(sms, d_out, d_cyc, json);
    run<LDS128>(sms, d_out, d_cyc, json);
    run<CREDUX>(sms, d_out, d_cyc, json);
    json += "  \"note\": \"8 independent chains/thread, 32 warps/SM, 2048 iterations; rates are SASS warp-instructions\"\n}\n";
    fputs(json.c_str(), stdout);
    return 0;
}
