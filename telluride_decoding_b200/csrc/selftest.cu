// Stand-alone diagnostic for libtdgram.so: runs the lagged Gram (tcgen05 and SIMT) on a list of
// shapes against a float64 host loop that applies the lag definition literally, and prints the
// relative Frobenius errors plus timings.  Usage:
//   selftest [cta_group=1|2|0] [case filter substring]
// A development tool (one gpurun = one table of results); the parity tests proper are in tests/.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "tdgram.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

struct Case {
  const char* name;
  int C, pre, post;
  int C2, pre2, post2;       // C2 = 0: single stream
  std::vector<int64_t> seg_len;
  std::vector<int64_t> seg_keep;   // empty = keep all
  bool host_check;
  std::vector<int64_t> seg_first;  // empty = 0
};

static void host_gram(const std::vector<float>& x, int C, int pre, int post, const std::vector<float>& y, int C2, int pre2,
                      int post2, const std::vector<int64_t>& start, const std::vector<int64_t>& len,
                      const std::vector<int64_t>& keep, std::vector<double>& G, const std::vector<int64_t>& first) {
  const int L1 = pre + 1 + post, L2 = C2 ? pre2 + 1 + post2 : 0;
  const int64_t D1 = (int64_t)L1 * C, D = D1 + (int64_t)L2 * C2;
  G.assign((size_t)D * D, 0.0);
  std::vector<double> row(D);
  for (size_t s = 0; s < len.size(); ++s) {
    for (int64_t t = first.empty() ? 0 : first[s]; t < keep[s]; ++t) {
      for (int k = 0; k < L1; ++k)
        for (int c = 0; c < C; ++c) {
          int64_t u = t - pre + k;
          row[(int64_t)k * C + c] = (u >= 0 && u < len[s]) ? (double)x[(start[s] + u) * C + c] : 0.0;
        }
      for (int k = 0; k < L2; ++k)
        for (int c = 0; c < C2; ++c) {
          int64_t u = t - pre2 + k;
          row[D1 + (int64_t)k * C2 + c] = (u >= 0 && u < len[s]) ? (double)y[(start[s] + u) * C2 + c] : 0.0;
        }
      for (int64_t i = 0; i < D; ++i) {
        const double ri = row[i];
        double* g = &G[(size_t)i * D];
        for (int64_t j = 0; j <= i; ++j) g[j] += ri * row[j];
      }
    }
  }
  for (int64_t i = 0; i < D; ++i)
    for (int64_t j = i + 1; j < D; ++j) G[(size_t)i * D + j] = G[(size_t)j * D + i];
}

static double rel_fro(const std::vector<float>& a, const std::vector<double>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { double d = (double)a[i] - b[i]; num += d * d; den += b[i] * b[i]; }
  return std::sqrt(num / (den > 0 ? den : 1));
}
static double rel_fro_ff(const std::vector<float>& a, const std::vector<float>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { double d = (double)a[i] - (double)b[i]; num += d * d; den += (double)b[i] * b[i]; }
  return std::sqrt(num / (den > 0 ? den : 1));
}

int main(int argc, char** argv) {
  setvbuf(stdout, nullptr, _IONBF, 0);
  int cta_group = argc > 1 ? atoi(argv[1]) : 1;
  const char* filter = argc > 2 ? argv[2] : "";
  int chunk_rows = argc > 3 ? atoi(argv[3]) : 0;
  int sms = 0, ccM = 0, ccm = 0;
  if (td_device_info(&sms, &ccM, &ccm)) { printf("device_info: %s\n", td_last_error()); return 2; }
  printf("device: %d SMs, sm_%d%d; cta_group=%d chunk_rows=%d\n", sms, ccM, ccm, cta_group, chunk_rows);

  std::vector<Case> cases = {
      {"c8_l1_n64", 8, 0, 0, 0, 0, 0, {64}, {}, true},
      {"c8_l1_n50", 8, 0, 0, 0, 0, 0, {50}, {}, true},
      {"c2_l3_n100", 2, 1, 1, 0, 0, 0, {100}, {}, true},
      {"c3_l4_2seg_keep", 3, 1, 2, 0, 0, 0, {333, 217}, {320, 200}, true},
      {"c3_l4_2seg_range", 3, 1, 2, 0, 0, 0, {333, 217}, {320, 200}, true, {37, 64}},
      {"c64_l4_n4096", 64, 0, 3, 0, 0, 0, {4096}, {}, true},
      {"c32_l33_n3000", 32, 16, 16, 0, 0, 0, {3000}, {}, true},
      {"c20_l5+c6_l3_cca", 20, 0, 4, 6, 1, 1, {1500, 901}, {1500, 890}, true},
      {"c64_l8_n20000", 64, 0, 7, 0, 0, 0, {20000}, {}, true},
      {"c64_l32_1h128Hz", 64, 0, 31, 0, 0, 0, std::vector<int64_t>(60, 7680), {}, false},
      {"c306_l10_n50000", 306, 0, 9, 0, 0, 0, {50000}, {}, false},
  };
  int failures = 0;
  for (const Case& cs : cases) {
    if (filter[0] && !strstr(cs.name, filter)) continue;
    const int n_seg = (int)cs.seg_len.size();
    std::vector<int64_t> start(n_seg), keep = cs.seg_keep.empty() ? cs.seg_len : cs.seg_keep;
    const std::vector<int64_t>& first = cs.seg_first;
    int64_t rows = 0;
    for (int s = 0; s < n_seg; ++s) { start[s] = rows; rows += cs.seg_len[s]; }
    printf("[case %s] begin\n", cs.name);
    std::mt19937 rng(1234);
    std::normal_distribution<float> nd(0.f, 1.f);
    std::vector<float> x((size_t)rows * cs.C), y((size_t)rows * std::max(cs.C2, 1));
    for (auto& v : x) v = nd(rng) + 0.25f;
    for (auto& v : y) v = nd(rng) - 0.5f;
    float *xd, *yd = nullptr;
    CK(cudaMalloc(&xd, x.size() * 4));
    CK(cudaMemcpy(xd, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    td_stream st[2];
    st[0] = {xd, cs.C, cs.C, cs.pre, cs.post, start.data(), cs.seg_len.data()};
    int ns = 1;
    int64_t D = (int64_t)cs.C * (cs.pre + 1 + cs.post);
    if (cs.C2) {
      CK(cudaMalloc(&yd, y.size() * 4));
      CK(cudaMemcpy(yd, y.data(), y.size() * 4, cudaMemcpyHostToDevice));
      st[1] = {yd, cs.C2, cs.C2, cs.pre2, cs.post2, start.data(), cs.seg_len.data()};
      ns = 2;
      D += (int64_t)cs.C2 * (cs.pre2 + 1 + cs.post2);
    }
    float *Gd, *Gs;
    CK(cudaMalloc(&Gd, (size_t)D * D * 4));
    CK(cudaMalloc(&Gs, (size_t)D * D * 4));
    td_gram_options opt = {};
    opt.method = 0;
    opt.cta_group = cta_group;
    opt.chunk_rows = chunk_rows;
    size_t wsb = 0;
    if (td_lagged_gram_workspace_bytes(st, ns, first.empty() ? nullptr : first.data(), keep.data(), n_seg, &opt, &wsb)) { printf("%s: ws: %s\n", cs.name, td_last_error()); return 2; }
    void* ws;
    CK(cudaMalloc(&ws, wsb ? wsb : 1));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms_umma = 0, ms_simt = 0;
    int rc = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      rc = td_lagged_gram_f32(st, ns, first.empty() ? nullptr : first.data(), keep.data(), n_seg, Gd, D, &opt, ws, wsb, nullptr);
      cudaEventRecord(e1);
      if (rc) { printf("%s: gram failed: %s\n", cs.name, td_last_error()); return 2; }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: kernel error: %s\n", cs.name, cudaGetErrorString(e)); return 3; }
      cudaEventElapsedTime(&ms_umma, e0, e1);
    }
    td_gram_options opt1 = {};
    opt1.method = 1;
    cudaEventRecord(e0);
    rc = td_lagged_gram_f32(st, ns, first.empty() ? nullptr : first.data(), keep.data(), n_seg, Gs, D, &opt1, nullptr, 0, nullptr);
    cudaEventRecord(e1);
    if (rc) { printf("%s: simt gram failed: %s\n", cs.name, td_last_error()); return 2; }
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms_simt, e0, e1);
    std::vector<float> g1((size_t)D * D), g2((size_t)D * D);
    CK(cudaMemcpy(g1.data(), Gd, g1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g2.data(), Gs, g2.size() * 4, cudaMemcpyDeviceToHost));
    double asym = 0;
    for (int64_t i = 0; i < D; ++i) for (int64_t j = 0; j < i; ++j) asym = std::max(asym, (double)std::fabs(g1[i * D + j] - g1[j * D + i]));
    int64_t frames = 0;
    for (size_t i = 0; i < keep.size(); ++i) frames += keep[i] - (first.empty() ? 0 : first[i]);
    double tflops = (double)frames * D * (D + 1) / (ms_umma * 1e-3) / 1e12;
    if (cs.host_check) {
      std::vector<double> G;
      host_gram(x, cs.C, cs.pre, cs.post, y, cs.C2, cs.pre2, cs.post2, start, cs.seg_len, keep, G, first);
      double e_umma = rel_fro(g1, G), e_simt = rel_fro(g2, G);
      bool ok = e_umma < 1e-5 && e_simt < 1e-5;
      if (!ok) ++failures;
      printf("%-22s D=%-6lld frames=%-8lld umma err %.3e  simt err %.3e  asym %.1e  umma %.3f ms (%.1f TF alg)  simt %.3f ms  %s\n", cs.name,
             (long long)D, (long long)frames, e_umma, e_simt, asym, ms_umma, tflops, ms_simt, ok ? "OK" : "FAIL");
      if (!ok && D <= 16) {
        for (int64_t i = 0; i < D; ++i) { for (int64_t j = 0; j < D; ++j) printf(" %9.3f/%9.3f", g1[i * D + j], G[i * D + j]); printf("\n"); }
      }
    } else {
      double e_rel = rel_fro_ff(g1, g2);
      bool ok = e_rel < 1e-5;
      if (!ok) ++failures;
      printf("%-22s D=%-6lld frames=%-8lld umma vs simt %.3e  asym %.1e  umma %.3f ms (%.1f TF alg)  simt %.3f ms  %s\n", cs.name, (long long)D,
             (long long)frames, e_rel, asym, ms_umma, tflops, ms_simt, ok ? "OK" : "FAIL");
    }
    fflush(stdout);
    cudaFree(xd); if (yd) cudaFree(yd); cudaFree(Gd); cudaFree(Gs); cudaFree(ws);
  }
  printf("selftest: %d failure(s)\n", failures);
  return failures ? 1 : 0;
}
