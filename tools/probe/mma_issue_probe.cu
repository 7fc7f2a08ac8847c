// mma_issue_probe — how does one thread's stream of tcgen05.mma.kind::i8 issue on sm_100a?
// Measures, for A taken from tensor memory (TS, what k_scan_mma uses) and from shared memory (SS), and for several
// N, the clock64 interval between consecutive issues from an EMPTY pipe and the time until the commit arrives.
// If issues return quickly until some depth and only then throttle to the execution rate, the hand-off work of an
// issuing thread can hide behind queued MMAs; if every issue takes an execution time, it cannot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_issue_probe mma_issue_probe.cu && ./mma_issue_probe
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t idesc_i8(int m, int n) {
  return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

constexpr int NMMA = 48;

// mode 0: TS, 1: SS.  gap: clock cycles of busy-waiting inserted after every 4th issue (an artificial hand-off).
__global__ void __launch_bounds__(128, 1) k_probe(int mode, int n_tile, int gap, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (n_tile * 32 + 128 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = idesc_i8(128, n_tile);
    // B: n_tile x 32 bytes, K-major core matrices: LBO = n_tile*16, SBO = 128.  A (SS): 128 x 32 bytes behind it.
    const uint64_t bdesc = kmajor_desc(smem_u32(smem), (uint32_t)n_tile * 16u, 128u);
    const uint64_t adesc = kmajor_desc(smem_u32(smem + n_tile * 32), 128u * 16u, 128u);
    const uint32_t a_tmem = tm + 2u * (uint32_t)n_tile;  // behind two accumulators, as in k_scan_mma
    // timestamps only every 4th issue (a clock read costs tens of cycles and would dominate a per-issue trace)
    long long t[NMMA + 2];
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < NMMA; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (mode == 0) mma_ts(tm, a_tmem, bdesc, idesc, (i + j) != 0);
        else mma_ss(tm, adesc, bdesc, idesc, (i + j) != 0);
      }
      t[i / 4 + 1] = clock64();
      if (gap > 0) {
        const long long until = t[i / 4 + 1] + gap;
        while (clock64() < until) {}
      }
    }
    t[NMMA] = t[NMMA / 4];
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done;
    do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done && clock64() - t[NMMA] < 20000000ll);  // bounded: a faulted MMA must not hang the box
    t[NMMA + 1] = clock64();
    if (blockIdx.x == 0)
      for (int i = 0; i < NMMA + 2; i++) out[i] = t[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}


// Probe 2: how long does a warp's tcgen05.st.32x32b of 4 KB (x32) / 8 KB (x64) take to issue, alone and while another
// thread keeps the tensor pipe busy with TS-mode MMAs on the same CTA?  (k_scan_mma's expansion warps feed the A
// operand this way.)
template <int X>
__device__ __forceinline__ void st_cols(uint32_t taddr, uint32_t v) {
  if (X == 32) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
  } else if (X == 64) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
  } else {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(v) : "memory");
  }
}

template <int X>
__global__ void __launch_bounds__(256, 1) k_probe_st(int with_mma, int wait_each, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = 208;
  for (int i = threadIdx.x; i < (n_tile * 32) / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  constexpr int NST = 64;
  if (warp == 0 && with_mma) {
    if (lane == 0) {
      const uint32_t idesc = idesc_i8(128, n_tile);
      const uint64_t bdesc = kmajor_desc(smem_u32(smem), (uint32_t)n_tile * 16u, 128u);
      int it = 0;
      while (!stop && it < 4000) {   // bounded
#pragma unroll
        for (int j = 0; j < 4; j++) mma_ts(tm, tm + 416u + 64u, bdesc, idesc, 1u);  // reads A columns 480..487
        it++;
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done;
      const long long t0 = clock64();
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      } while (!done && clock64() - t0 < 20000000ll);
    }
  } else if (warp >= 4) {
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 416u;  // A area: columns 416..479
    long long t0 = clock64();
    long long acc_issue = 0;
    for (int i = 0; i < NST; i++) {
      const long long a = clock64();
      st_cols<X>(lane_addr + (X == 64 ? 0u : (uint32_t)(i & 1) * 32u), (uint32_t)i);
      acc_issue += clock64() - a;
      if (wait_each) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) {
      out[(warp - 4) * 2] = t1 - t0;
      out[(warp - 4) * 2 + 1] = acc_issue;
    }
    __syncwarp();
    if (warp == 4 && lane == 0) stop = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

template <int X>
static void run_st(long long* d) {
  const int smem = 208 * 32 + 1024;
  for (int with_mma = 0; with_mma < 2; with_mma++)
    for (int wait_each = 0; wait_each < 2; wait_each++) {
      long long h[8];
      for (int rep = 0; rep < 2; rep++) {
        k_probe_st<X><<<148, 256, smem>>>(with_mma, wait_each, d);
        CK(cudaDeviceSynchronize());
      }
      CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
      printf("tcgen05.st.32x32b.x%d (%d B/warp) mma=%d wait_each=%d: per store: total %.1f cyc, issue %.1f cyc (warp 4); warps 5-7 total %.1f %.1f %.1f\n",
             X, X * 128, with_mma, wait_each, h[0] / 64.0, h[1] / 64.0, h[2] / 64.0, h[4] / 64.0, h[6] / 64.0);
    }
}

// Probe 3: the reverse question — how fast do N=208 TS-mode MMAs run while other warps of the CTA stream
// tcgen05.st (the A operand writers) and / or tcgen05.ld (the epilogue readers) through tensor memory?
__global__ void __launch_bounds__(512, 1) k_probe_contention(int n_tile, int st_warps, int ld_warps, int st_pause, int ld_pause, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (n_tile * 32) / 4; i += 512) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_ptr;
  const uint32_t a_col = 2u * (uint32_t)n_tile;
  long long nst = 0, nld = 0;
  if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = idesc_i8(128, n_tile);
      const uint64_t bdesc = kmajor_desc(smem_u32(smem), (uint32_t)n_tile * 16u, 128u);
      constexpr int N3 = 512;
      const long long t0 = clock64();
#pragma unroll 1
      for (int i = 0; i < N3; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) mma_ts(tm, tm + a_col + 64u + (uint32_t)j * 8u, bdesc, idesc, 1u);  // A: third stage
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      } while (!done && clock64() - t1 < 20000000ll);
      const long long t2 = clock64();
      stop = 1;
      if (blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
        out[2] = N3;
      }
    }
  } else if (warp >= 4 && warp < 4 + st_warps) {
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + a_col;  // stages 0/1 (the MMAs read stage 2)
    int i = 0;
    while (!stop && i < 200000) {
      st_cols<32>(lane_addr + (uint32_t)(i & 1) * 32u, (uint32_t)i);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (st_pause) { const long long u = clock64() + st_pause; while (clock64() < u) {} }
      i++;
    }
    nst = i;
  } else if (warp >= 8 && warp < 8 + ld_warps) {
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)n_tile;  // second accumulator
    int i = 0, sink = 0;
    while (!stop && i < 200000) {
      int r[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                     "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(lane_addr + (uint32_t)((i % 12) * 16)) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      sink += r[0] + r[15];
      if (ld_pause) { const long long u = clock64() + ld_pause; while (clock64() < u) {} }
      i++;
    }
    nld = i + (sink == 123456789);
  }
  if (blockIdx.x == 0 && lane == 0) {
    if (warp == 4) out[3] = nst;
    if (warp == 8) out[4] = nld;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

static void run_contention(long long* d) {
  const int smem = 208 * 32 + 1024;
  const int cfg[][4] = {{0, 0, 0, 0}, {4, 0, 0, 0}, {0, 8, 0, 0}, {4, 8, 0, 0}, {4, 0, 300, 0}, {0, 8, 0, 300}, {4, 8, 300, 300}, {0, 4, 0, 0}, {0, 8, 0, 1000}};
  for (auto& c : cfg) {
    long long h[8] = {0};
    CK(cudaMemset(d, 0, sizeof(h)));
    for (int rep = 0; rep < 2; rep++) {
      k_probe_contention<<<148, 512, smem>>>(208, c[0], c[1], c[2], c[3], d);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    printf("contention: N=208 TS, st warps %d (pause %d), ld warps %d (pause %d): %.1f cyc per MMA (issue-side %.1f); during that one st warp did %lld stores (4 KB), one ld warp %lld loads (2 KB)\n",
           c[0], c[2], c[1], c[3], (double)h[1] / h[2], (double)h[0] / h[2], h[3], h[4]);
  }
}

int main(int argc, char** argv) {
  long long* d;
  CK(cudaMalloc(&d, (NMMA + 2) * sizeof(long long)));
  if (argc > 1 && !strcmp(argv[1], "peak")) {
    // bench.py's roofline denominator: the tcgen05.mma.kind::i8 rate of an otherwise idle SM (M=128, N=208, K=32,
    // A in tensor memory — the shape k_scan_mma issues), 512 back-to-back MMAs per SM on all 148 SMs, in SM cycles
    const int smem = 208 * 32 + 1024;
    long long h[8] = {0};
    for (int rep = 0; rep < 3; rep++) {
      k_probe_contention<<<148, 512, smem>>>(208, 0, 0, 0, 0, d);
      CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    const double cyc = (double)h[1] / (double)h[2];
    printf("{\"cycles_per_mma\": %.3f, \"m\": 128, \"n\": 208, \"k\": 32, \"int8_ops_per_cycle_per_sm\": %.1f, \"sms\": 148}\n",
           cyc, 2.0 * 128 * 208 * 32 / cyc);
    return 0;
  }
  const int smem = 256 * 32 + 128 * 32 + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int ns[] = {112, 208, 256};
  for (int mode = 0; mode < 2; mode++)
    for (int n : ns)
      for (int gap : {0, 200}) {
        if (mode == 0 && n == 256) continue;  // TS needs columns behind two accumulators
        long long h[NMMA + 2];
        printf("%s N=%d gap=%d\n", mode ? "SS" : "TS", n, gap); fflush(stdout);
        for (int rep = 0; rep < 2; rep++) {   // second run: warm instruction cache
          k_probe<<<148, 128, smem>>>(mode, n, gap, d);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
        printf("%s N=%d gap=%d: issue intervals:", mode ? "SS" : "TS", n, gap);
        for (int i = 0; i < NMMA / 4; i++) printf(" %lld", h[i + 1] - h[i]);
        printf(" (per group of 4)");
        printf(" | issue total %lld, until commit %lld, per MMA %.1f\n", h[NMMA] - h[0], h[NMMA + 1] - h[0],
               (double)(h[NMMA + 1] - h[0]) / NMMA);
      }
  run_contention(d);
  run_st<16>(d);
  run_st<32>(d);
  run_st<64>(d);
  return 0;
}
