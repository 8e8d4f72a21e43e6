// potential.cu -- gravitational potential on a set of points and per-particle energies
// (reference wendy/wendy.py:494-517 `potential`, wendy/wendy.py:466-470 `energy(individual=True)`).
//
// The reference evaluates sum_i m_i |x_i - y_j| by an O(N*Y) numpy broadcast.  Here: one radix sort of
// the positions, exclusive prefix sums M_k = sum_{i<k} m_i and S_k = sum_{i<k} m_i x_i in sorted order,
// and per query a binary search:
//     Phi(y) = y (2 M_k - M_tot) + (S_tot - 2 S_k),   k = #{x_i < y}
// i.e. O((N + Y) log N).  Sums are fp64 in a FIXED tree order (8 consecutive elements per thread, warp
// scan, warp totals, tile totals scanned by one CTA), so results are reproducible run to run; they agree
// with the reference's pairwise numpy sums to rounding (a few ulp of max|Phi|), not bit for bit.
#include <string>

#include "common.cuh"
#include "internal.h"

namespace wendy {

namespace {
constexpr int PT = 256;         // threads per CTA
constexpr int PE = 8;           // consecutive sorted elements per thread
constexpr int PTILE = PT * PE;  // 2048 elements per CTA

__global__ void pot_keys_kernel(const double *__restrict__ x, long long n, uint64_t *__restrict__ key,
                                uint32_t *__restrict__ val) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    key[i] = key_from_double(x[i]);
    val[i] = (uint32_t)i;
  }
}

// exclusive scan of (a, b) over the CTA in thread order; returns the CTA totals in ta, tb
__device__ __forceinline__ void block_exclusive_scan2(double &a, double &b, double &ta, double &tb,
                                                      double (*sw)[PT / 32 + 1]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double ia = a, ib = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double ua = __shfl_up_sync(WENDY_FULL_MASK, ia, o), ub = __shfl_up_sync(WENDY_FULL_MASK, ib, o);
    if (lane >= o) {
      ia += ua;
      ib += ub;
    }
  }
  if (lane == 31) {
    sw[0][wid] = ia;
    sw[1][wid] = ib;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ra = 0., rb = 0.;
    for (int w = 0; w < PT / 32; w++) {
      const double ca = sw[0][w], cb = sw[1][w];
      sw[0][w] = ra;
      sw[1][w] = rb;
      ra += ca;
      rb += cb;
    }
    sw[0][PT / 32] = ra;
    sw[1][PT / 32] = rb;
  }
  __syncthreads();
  a = (ia - a) + sw[0][wid];
  b = (ib - b) + sw[1][wid];
  ta = sw[0][PT / 32];
  tb = sw[1][PT / 32];
}

// pass 1 (WRITE = 0): tile totals; pass 2 (WRITE = 1): exclusive prefixes per sorted element
template <int WRITE>
__global__ void __launch_bounds__(PT)
pot_prefix_kernel(const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sval,
                  const double *__restrict__ m, long long n, double *__restrict__ xs, double2 *__restrict__ tiles,
                  double *__restrict__ Mex, double *__restrict__ Sex) {
  __shared__ double sw[2][PT / 32 + 1];
  const long long k0 = (long long)blockIdx.x * PTILE + (long long)threadIdx.x * PE;
  double pm[PE], ps[PE], a = 0., b = 0.;
#pragma unroll
  for (int q = 0; q < PE; q++) {
    pm[q] = a;
    ps[q] = b;
    if (k0 + q < n) {
      const double xx = double_from_key(skey[k0 + q]);
      const double mm = m[sval[k0 + q]];
      if (!WRITE) xs[k0 + q] = xx;
      a += mm;
      b += mm * xx;
    }
  }
  double ta, tb;
  block_exclusive_scan2(a, b, ta, tb, sw);
  if (!WRITE) {
    if (threadIdx.x == 0) tiles[blockIdx.x] = make_double2(ta, tb);
  } else {
    const double2 base = tiles[blockIdx.x];
#pragma unroll
    for (int q = 0; q < PE; q++)
      if (k0 + q < n) {
        Mex[k0 + q] = base.x + (a + pm[q]);
        Sex[k0 + q] = base.y + (b + ps[q]);
      }
  }
}

// one CTA: exclusive scan of the tile totals in place; tot[0..1] = (M_tot, S_tot)
__global__ void __launch_bounds__(PT) pot_scan_tiles_kernel(double2 *__restrict__ tiles, int nt, double *__restrict__ tot) {
  __shared__ double sw[2][PT / 32 + 1];
  const int per = (nt + PT - 1) / PT;
  const int i0 = threadIdx.x * per;
  double a = 0., b = 0.;
  for (int i = i0; i < min(nt, i0 + per); i++) {
    a += tiles[i].x;
    b += tiles[i].y;
  }
  double ta, tb;
  block_exclusive_scan2(a, b, ta, tb, sw);
  for (int i = i0; i < min(nt, i0 + per); i++) {
    const double2 c = tiles[i];
    tiles[i] = make_double2(a, b);
    a += c.x;
    b += c.y;
  }
  if (threadIdx.x == 0) {
    tot[0] = ta;
    tot[1] = tb;
  }
}

__device__ __forceinline__ double phi_at(double q, double Mk, double Sk, double Mt, double St) {
  return q * (2. * Mk - Mt) + (St - 2. * Sk);
}

__global__ void pot_query_kernel(const double *__restrict__ y, long long Y, const double *__restrict__ xs,
                                 const double *__restrict__ Mex, const double *__restrict__ Sex,
                                 const double *__restrict__ tot, long long n, double twopiG, double omega2,
                                 double *__restrict__ out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= Y) return;
  const double q = y[j];
  long long lo = 0, hi = n;  // first k with xs[k] >= q
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (xs[mid] < q) lo = mid + 1;
    else hi = mid;
  }
  const double Mt = tot[0], St = tot[1];
  const double Mk = lo < n ? Mex[lo] : Mt, Sk = lo < n ? Sex[lo] : St;
  double r = twopiG * phi_at(q, Mk, Sk, Mt, St);
  if (omega2 >= 0.) r += omega2 * q * q / 2.;
  out[j] = r;
}

__global__ void pot_individual_kernel(const double *__restrict__ xs, const uint32_t *__restrict__ sval,
                                      const double *__restrict__ v, const double *__restrict__ m,
                                      const double *__restrict__ Mex, const double *__restrict__ Sex,
                                      const double *__restrict__ tot, long long n, double twopiG, double omega2,
                                      double *__restrict__ out) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t id = sval[k];
  const double q = xs[k], mk = m[id], vv = v[id];
  double r = twopiG * mk * phi_at(q, Mex[k], Sex[k], tot[0], tot[1]) + mk * vv * vv / 2.;
  if (omega2 >= 0.) r += mk * omega2 * q * q / 2.;
  out[id] = r;
}

// a host-or-device input made available on the device
struct DevArray {
  void *p = nullptr;
  bool owned = false;
  cudaError_t get(const void *src, size_t bytes, cudaStream_t st) {
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, src);
    if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
      p = const_cast<void *>(src);
      return cudaSuccess;
    }
    cudaGetLastError();
    e = cudaMalloc(&p, bytes ? bytes : 8);
    if (e != cudaSuccess) return e;
    owned = true;
    return cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, st);
  }
  ~DevArray() {
    if (owned && p) cudaFree(p);
  }
};
}  // namespace

#define PCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e__); rc = -1; goto done; } } while (0)

// individual == 0: out[Y] = potential at y (v unused); individual == 1: out[N] = per-particle energies.
// Every array argument may be a HOST or a DEVICE pointer.
int potential_eval(cudaStream_t st, const double *x, const double *v, const double *m, long long N,
                   const double *y, long long Y, double twopiG, double omega2, double *out, int individual,
                   std::string &err) {
  int rc = 0;
  const size_t n = (size_t)N;
  const int nt = (int)((n + PTILE - 1) / PTILE);
  const long long n_out = individual ? N : Y;
  RadixScratch rs;
  DevArray dx, dv, dm, dy;
  double *xs = nullptr, *Mex = nullptr, *Sex = nullptr, *tot = nullptr, *dout = nullptr;
  double2 *tiles = nullptr;
  bool out_dev = false;
  int res = 0;
  {
    cudaPointerAttributes at;
    out_dev = cudaPointerGetAttributes(&at, out) == cudaSuccess &&
              (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    cudaGetLastError();
  }
  PCK(dx.get(x, n * sizeof(double), st));
  PCK(dm.get(m, n * sizeof(double), st));
  if (individual) PCK(dv.get(v, n * sizeof(double), st));
  else PCK(dy.get(y, (size_t)Y * sizeof(double), st));
  for (int i = 0; i < 2; i++) {
    PCK(cudaMalloc(&rs.key[i], n * sizeof(uint64_t)));
    PCK(cudaMalloc(&rs.val[i], n * sizeof(uint32_t)));
  }
  PCK(cudaMalloc(&rs.table, radix_table_entries(n) * sizeof(uint32_t)));
  PCK(cudaMalloc(&rs.sums, (radix_sums_entries(n) + 1) * sizeof(uint32_t)));
  rs.n_alloc = n;
  PCK(cudaMalloc(&xs, n * sizeof(double)));
  PCK(cudaMalloc(&Mex, n * sizeof(double)));
  PCK(cudaMalloc(&Sex, n * sizeof(double)));
  PCK(cudaMalloc(&tiles, (size_t)nt * sizeof(double2)));
  PCK(cudaMalloc(&tot, 2 * sizeof(double)));
  if (out_dev) dout = out;
  else PCK(cudaMalloc(&dout, (size_t)(n_out ? n_out : 1) * sizeof(double)));

  pot_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const double *)dx.p, N, rs.key[0], rs.val[0]);
  res = radix_sort_pairs(st, rs, n, 0, 1u);
  pot_prefix_kernel<0><<<nt, PT, 0, st>>>(rs.key[res], rs.val[res], (const double *)dm.p, N, xs, tiles, Mex, Sex);
  pot_scan_tiles_kernel<<<1, PT, 0, st>>>(tiles, nt, tot);
  pot_prefix_kernel<1><<<nt, PT, 0, st>>>(rs.key[res], rs.val[res], (const double *)dm.p, N, xs, tiles, Mex, Sex);
  if (individual) {
    pot_individual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        xs, rs.val[res], (const double *)dv.p, (const double *)dm.p, Mex, Sex, tot, N, twopiG, omega2, dout);
  } else if (Y > 0) {
    pot_query_kernel<<<(unsigned)((Y + 255) / 256), 256, 0, st>>>((const double *)dy.p, Y, xs, Mex, Sex, tot, N,
                                                                   twopiG, omega2, dout);
  }
  PCK(cudaGetLastError());
  if (!out_dev && n_out > 0)
    PCK(cudaMemcpyAsync(out, dout, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
  PCK(cudaStreamSynchronize(st));
done:
  for (int i = 0; i < 2; i++) {
    cudaFree(rs.key[i]);
    cudaFree(rs.val[i]);
  }
  cudaFree(rs.table);
  cudaFree(rs.sums);
  cudaFree(xs);
  cudaFree(Mex);
  cudaFree(Sex);
  cudaFree(tiles);
  cudaFree(tot);
  if (!out_dev) cudaFree(dout);
  return rc;
}

}  // namespace wendy
