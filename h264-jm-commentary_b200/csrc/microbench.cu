// microbench.cu — measured issue rates of the integer instructions the ME kernels are made of.
//
// MEASURED_PEAKS.json carries only HBM and bf16 numbers; the roofline denominator of the search
// kernel (SURVEY.md §8(d) "INT peak") is measured here.  Every test runs CH independent dependent
// chains per thread of one SASS instruction (or of the search kernel's per-candidate mix), 1024
// threads per SM x all SMs, and is reported three ways that must agree with each other:
//
//   ms            CUDA events around the last launches of a >= 150 ms back-to-back run (steady clocks:
//                 a full-issue integer kernel runs into the 1000 W power cap like a dense GEMM does,
//                 and the SM clock settles well below clocks.max.sm — that IS the sustained peak)
//   cycles        in-kernel clock64 of every CTA, first instruction to the barrier behind its last warp; the slowest
//                 CTA spans the launch  ->  mhz_clock64 = max cycles / ms
//   mhz_nvml      SM clock sampled through NVML by a host thread while the run is in flight (median)
//
// warp_instr_per_clk_per_sm counts SASS instructions (ptxas fuses two dependent min.u32 into one
// VIMNMX3: the per-test SASS counts below were read off `cuobjdump -sass microbench`), so it can
// never exceed the 4 schedulers of an SM; lane-op rates count ALGORITHMIC operations (a fused
// VIMNMX3 is two mins), the unit of roofline.achieved in bench.py.
// Output: one JSON object on stdout (profiles/INT_PEAKS_r02.json is a copy of one run).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CHECK(x)                                                                  \
    do {                                                                          \
        cudaError_t e = (x);                                                      \
        if (e != cudaSuccess) {                                                   \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));               \
            exit(1);                                                              \
        }                                                                         \
    } while (0)

constexpr int ITERS = 8192;
constexpr int CH = 8;

enum Op { VSAD, IADD3, IMAD, LEA, VIMNMX, VIMNMX3, LOP3, PRMT, SHF, MIX_SAD_IMAD, MIX_SAD_IMAD_MIN, MIX_ME, LDS32, LDS128,
          CREDUX, N_OPS };
const char *kNames[N_OPS] = {"vabsdiff4_acc", "iadd3", "imad", "lea", "vimnmx_u32", "vimnmx3_u32", "lop3", "prmt", "shf",
                             "mix_2sad_1imad", "mix_64sad_66imad_41min", "mix_me_171", "lds32", "lds128", "credux_min"};
// per thread and loop iteration: algorithmic operations, and SASS instructions they become
//   mix_64sad_66imad_41min: 64 VABSDIFF4 + 64 IMAD + 40 min.u32 (ptxas: 16 VIMNMX3 + 8 VIMNMX)
//   mix_me_171:             64 VABSDIFF4 + 64 IMAD + 80 min.u32 (40 VIMNMX3) — two candidates' minima per SAD set
const double kAlgOps[N_OPS] = {CH, CH, CH, CH, CH, 2 * CH, CH, CH, CH, CH * 1.5, CH * 21, CH * 26, CH, CH, CH};
const double kSassOps[N_OPS] = {CH, CH, CH, CH, CH, CH, CH, CH, CH, CH * 1.5, CH * 19, CH * 21, CH, CH, CH};

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
            // the per-candidate instruction mix of the search kernel: 64 SAD, 25 adds + 41 packs, 41 mins
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
    __syncthreads();                       // every warp of the CTA is done: the schedulers favour the older warps, so
    long long t1 = clock64();              // warp 0's own loop time alone under-reports the time the SM was busy
    unsigned s = 0;
    for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ---- NVML through dlopen (the library exists on the GPU box only) ------------------------------------------
struct Nvml {
    void *h = nullptr, *dev = nullptr;
    int (*clock)(void *, int, unsigned *) = nullptr;
    int (*power)(void *, unsigned *) = nullptr;
    bool ok = false;
    void open(int cuda_dev)
    {
        h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!h) return;
        auto init = (int (*)())dlsym(h, "nvmlInit_v2");
        auto by_pci = (int (*)(const char *, void **))dlsym(h, "nvmlDeviceGetHandleByPciBusId_v2");
        clock = (int (*)(void *, int, unsigned *))dlsym(h, "nvmlDeviceGetClockInfo");
        power = (int (*)(void *, unsigned *))dlsym(h, "nvmlDeviceGetPowerUsage");
        char bus[32];
        if (!init || !by_pci || !clock || init() != 0) return;
        if (cudaDeviceGetPCIBusId(bus, sizeof bus, cuda_dev) != cudaSuccess) return;
        ok = by_pci(bus, &dev) == 0;
    }
};

struct Sampler {
    Nvml &n;
    std::atomic<bool> stop{false};
    std::vector<unsigned> mhz, mw;
    std::thread th;
    explicit Sampler(Nvml &nv) : n(nv)
    {
        if (!n.ok) return;
        th = std::thread([this] {
            while (!stop.load()) {
                unsigned c = 0, p = 0;
                if (n.clock(n.dev, 1 /* NVML_CLOCK_SM */, &c) == 0) mhz.push_back(c);
                if (n.power && n.power(n.dev, &p) == 0) mw.push_back(p);
                std::this_thread::sleep_for(std::chrono::milliseconds(3));
            }
        });
    }
    void finish(double &mhz_med, double &watts_max)
    {
        stop = true;
        if (th.joinable()) th.join();
        mhz_med = 0; watts_max = 0;
        if (!mhz.empty()) {
            // the second half of the run: the clock has settled under the load by then
            std::vector<unsigned> v(mhz.begin() + mhz.size() / 2, mhz.end());
            std::sort(v.begin(), v.end());
            mhz_med = v[v.size() / 2];
        }
        for (unsigned p : mw) watts_max = std::max(watts_max, p * 1e-3);
    }
};

template <int OP>
void run(int sms, unsigned *d_out, long long *d_cyc, Nvml &nvml, double run_ms, std::string &json)
{
    const int blocks = sms * 4, threads = 256;   // 4 CTAs x 8 warps = 32 warps per SM, 8 per SMSP
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    // calibrate one launch, then run back to back for >= run_ms with the sampler on; the events bracket the last third
    CHECK(cudaEventRecord(e0));
    k<OP><<<blocks, threads>>>(d_out, d_cyc, 1);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms1 = 0;
    CHECK(cudaEventElapsedTime(&ms1, e0, e1));
    const int n = std::max(6, (int)(run_ms / std::max(ms1, 1e-3f))), n_timed = std::max(2, n / 3);
    Sampler smp(nvml);
    for (int i = 0; i < n - n_timed; i++) k<OP><<<blocks, threads>>>(d_out, d_cyc, 1);
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < n_timed; i++) k<OP><<<blocks, threads>>>(d_out, d_cyc, 1);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    double mhz_nvml = 0, watts = 0;
    smp.finish(mhz_nvml, watts);
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= n_timed;
    std::vector<long long> cyc(blocks);
    CHECK(cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double avg = 0, mx = 0;
    for (long long v : cyc) { avg += (double)v; mx = std::max(mx, (double)v); }
    avg /= blocks;
    const double sass_per_sm = kSassOps[OP] * ITERS * 32.0;       // SASS warp-instructions per SM (32 warps)
    const double ipc_clock64 = sass_per_sm / mx;
    // the slowest CTA spans the kernel (the 4 CTAs of an SM share its schedulers): its cycles over the event time of
    // one launch is the SM clock if launches follow each other without a gap
    const double mhz_clock64 = mx / (ms * 1e-3) * 1e-6;
    const double ipc_nvml = mhz_nvml > 0 ? sass_per_sm / (ms * 1e-3 * mhz_nvml * 1e6) : 0;
    const double lane_ops_s = kAlgOps[OP] * ITERS * (double)blocks * threads / (ms * 1e-3);
    char buf[768];
    snprintf(buf, sizeof buf,
             "  \"%s\": {\"tera_lane_ops_per_s\": %.3f, \"ms\": %.4f, \"launches\": %d, \"first_launch_ms\": %.4f, "
             "\"sass_warp_instr_per_clk_per_sm\": %.3f, \"sass_warp_instr_per_clk_per_sm_nvml_clock\": %.3f, "
             "\"avg_cycles\": %.0f, \"max_cycles\": %.0f, \"mhz_clock64\": %.0f, \"mhz_nvml\": %.0f, \"power_w_max\": %.0f, "
             "\"alg_ops_per_iter\": %.1f, \"sass_ops_per_iter\": %.1f},\n",
             kNames[OP], lane_ops_s * 1e-12, ms, n, ms1, ipc_clock64, ipc_nvml, avg, mx, mhz_clock64, mhz_nvml, watts,
             kAlgOps[OP], kSassOps[OP]);
    json += buf;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main(int argc, char **argv)
{
    // microbench [run_ms] [mix-only]: run_ms of back-to-back launches per test (default 150)
    const double run_ms = argc > 1 ? atof(argv[1]) : 150.0;
    const bool mix_only = argc > 2 && !strcmp(argv[2], "mix-only");
    int dev = 0;
    CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CHECK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    Nvml nvml;
    nvml.open(dev);
    unsigned *d_out;
    long long *d_cyc;
    CHECK(cudaMalloc(&d_out, sizeof(unsigned) * sms * 4 * 256));
    CHECK(cudaMalloc(&d_cyc, sizeof(long long) * sms * 4));
    std::string json = "{\n";
    char head[512];
    snprintf(head, sizeof head, "  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"nvml\": %s, \"run_ms_per_test\": %.0f,\n",
             prop.name, sms, prop.clockRate, nvml.ok ? "true" : "false", run_ms);
    json += head;
    if (!mix_only) {
        run<VSAD>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<IADD3>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<IMAD>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<LEA>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<VIMNMX>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<VIMNMX3>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<LOP3>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<PRMT>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<SHF>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<MIX_SAD_IMAD>(sms, d_out, d_cyc, nvml, run_ms, json);
    }
    run<MIX_SAD_IMAD_MIN>(sms, d_out, d_cyc, nvml, run_ms, json);
    run<MIX_ME>(sms, d_out, d_cyc, nvml, run_ms, json);
    if (!mix_only) {
        run<LDS32>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<LDS128>(sms, d_out, d_cyc, nvml, run_ms, json);
        run<CREDUX>(sms, d_out, d_cyc, nvml, run_ms, json);
    }
    json += "  \"note\": \"8 independent chains/thread, 32 warps/SM, 8192 iterations per launch; launches back to back for run_ms, "
            "events around the last third; IPC counts SASS instructions, lane-op rates count algorithmic operations\"\n}\n";
    fputs(json.c_str(), stdout);
    return 0;
}
