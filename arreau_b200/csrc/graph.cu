// K1 -- periodic radius graph over the 27 lattice images (sm_100a).
//
// Replaces diffusion/diffusion_helpers.py:328-564 (radius_graph_pbc).  The reference materialises
// ~10 tensors of 27*sum(n_g^2) rows; here one warp owns one receiver atom and walks its 27*n_g
// candidates in (j, cell) order -- lane = image cell (offset in registers), loop over the senders j
// (one broadcast load of pos_j) -- filters them with a ballot, and (under the neighbour cap) keeps
// the `cap` nearest through a small shared-memory selection buffer, pre-filtered by a log-spaced
// distance-bin bound that the count pass derives from a per-receiver histogram.  No candidate list ever reaches HBM: the only traffic is pos/lattice in
// (L1/L2 resident per crystal) and the surviving edges out.
//
// Arithmetic is fp64 with explicit round-to-nearest adds/muls (no FMA contraction) in the
// reference's operation order, so thresholds and the cap order are decided on the same bits as
// the reference:   off_k = (c0*a + c1*b) + c2*c ;  dir = (pos_j + off_k) - pos_i ;
//                  d2 = (dx*dx + dy*dy) + dz*dz          (helpers:390-409)
// The image loop is only +-1 cells on unwrapped positions (quirk B3), so there is nothing to bin:
// with n_g <= 236 atoms a cell list would cost more than the 27*n_g distance tests it saves
// (C3: 276 M candidates = 5.5 GDFLOP, ~0.2 ms of fp64 on B200).
#include "common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kSelBuf = 256;      // selection buffer entries per warp
constexpr int kMaxCap = kSelBuf - 32;

__device__ __forceinline__ void cell_of(int k, int& c0, int& c1, int& c2) {
  // k-th element of itertools.product((-1,0,1), repeat=3)   (helpers:10)
  c0 = k / 9 - 1;
  c1 = (k / 3) % 3 - 1;
  c2 = k % 3 - 1;
}

// Lane = image k (27 of 32 lanes), loop over the senders j: the image offset lives in registers, pos_j is one
// broadcast load, and (iteration j, lane k) is the reference's candidate order c = 27 j + k.
__device__ __forceinline__ void lane_offset(const double* __restrict__ lat, int lane, double (&off)[3]) {
  int c0, c1, c2;
  cell_of(lane < 27 ? lane : 0, c0, c1, c2);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    // bmm(lattice^T, cells): sum over lattice rows m = 0,1,2 in order (helpers:390-393)
    const double t0 = __dmul_rn((double)c0, lat[0 * 3 + a]);
    const double t1 = __dmul_rn((double)c1, lat[1 * 3 + a]);
    const double t2 = __dmul_rn((double)c2, lat[2 * 3 + a]);
    off[a] = __dadd_rn(__dadd_rn(t0, t1), t2);
  }
}
__device__ __forceinline__ double lane_d2(const double* __restrict__ pj, const double (&off)[3], double pix, double piy,
                                          double piz, double& dx, double& dy, double& dz) {
  dx = __dadd_rn(__dadd_rn(pj[0], off[0]), -pix);
  dy = __dadd_rn(__dadd_rn(pj[1], off[1]), -piy);
  dz = __dadd_rn(__dadd_rn(pj[2], off[2]), -piz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Image culling.  A candidate (j, cell c) can only be in range if, along every lattice axis a, the fractional
// displacement f_j[a] + c[a] - f_i[a] is at most r * |b_a| in magnitude (b_a = the reciprocal vector of axis a: the
// projection of a Cartesian vector of length <= r on b_a cannot exceed that).  With [lo, hi] the fractional bounding box
// of the crystal's atoms, a whole image cell is therefore skipped when its shifted box misses that window on one axis
// (a conservative, slightly widened test: the exact fp64 distance test below still decides every surviving candidate,
// so the edge list is unchanged).  The A surviving images of a receiver are dealt to the lanes in ascending cell
// order, floor(32 / A) senders at a time, which keeps the reference's (j, cell) candidate order and cuts the walk from
// n to n / floor(32 / A) iterations: A is ~7 of 27 in a 15 A cell at r = 7 A, ~9 in a 9 A cell at r = 5 A, and 27 in the
// ~1 A cells of the sampler's first steps.  Used for crystals of at least kCullMinAtoms atoms (measured at 40 atoms per
// crystal: the box pass costs more than the shorter walk saves, 0.35 vs 0.27 ms per step).  The culled walk is a SEPARATE
// template instantiation of both kernels (its per-lane (image, sender) bookkeeping costs ~20 registers, which slowed the
// plain walk of small cells when both lived in one kernel, scratch/attic/README.md); the launchers pick it when the
// batch averages >= kCullMinAtoms atoms per crystal (C3: 200-atom supercells, 1.28 -> 0.82 ms per step at 15 A cells).
constexpr int kCullMinAtoms = 64;
struct LaneImage {
  int img;      // this lane's image cell (0..26)
  int jj;       // this lane's sender within a batch of q
  int q;        // senders per iteration
  bool live;
};
template <bool kCull>
__device__ __forceinline__ LaneImage active_images(const double* __restrict__ lat, const double* __restrict__ pos, int start,
                                                   int n, double pix, double piy, double piz, double r2, int lane) {
  if constexpr (!kCull) {      // the plain walk: lane = image cell, one sender per iteration (compile-time constants)
    LaneImage li;
    li.img = lane < 27 ? lane : 0;
    li.jj = 0;
    li.q = 1;
    li.live = lane < 27;
    return li;
  }
  const double ax = lat[0], ay = lat[1], az = lat[2], bx = lat[3], by = lat[4], bz = lat[5], cx = lat[6], cy = lat[7], cz = lat[8];
  // reciprocal vectors (rows): f[a] = pos . rec[a]
  double rec[3][3] = {{by * cz - bz * cy, bz * cx - bx * cz, bx * cy - by * cx},
                      {cy * az - cz * ay, cz * ax - cx * az, cx * ay - cy * ax},
                      {ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx}};
  const double det = ax * rec[0][0] + ay * rec[0][1] + az * rec[0][2];
  unsigned mask = 0x07ffffffu;
  // small crystals: the box pass below (n / 32 iterations plus the reductions) costs as much as the walk it could save
  if (n >= kCullMinAtoms && isfinite(det) && fabs(det) > 1e-200) {
    const double inv = 1.0 / det, r = sqrt(r2);
    double tol[3], lo[3], hi[3], fi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int d = 0; d < 3; ++d) rec[a][d] *= inv;
      tol[a] = r * sqrt(rec[a][0] * rec[a][0] + rec[a][1] * rec[a][1] + rec[a][2] * rec[a][2]) * (1.0 + 1e-6) + 1e-9;
      fi[a] = pix * rec[a][0] + piy * rec[a][1] + piz * rec[a][2];
      lo[a] = INFINITY;
      hi[a] = -INFINITY;
    }
    for (int j = lane; j < n; j += 32) {
      const double* pj = pos + 3 * (size_t)(start + j);
      const double x = pj[0], y = pj[1], z = pj[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double f = x * rec[a][0] + y * rec[a][1] + z * rec[a][2];
        lo[a] = fmin(lo[a], f);
        hi[a] = fmax(hi[a], f);
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
        hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
      }
    bool ok = lane < 27;
    if (ok) {
      int c[3];
      cell_of(lane, c[0], c[1], c[2]);
      bool sane = true, in = true;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const double w = tol[a] + 1e-6 * (fabs(lo[a]) + fabs(hi[a]) + fabs(fi[a]));     // rounding of the projections
        sane = sane && isfinite(w);
        in = in && (lo[a] + c[a] - fi[a] <= w) && (hi[a] + c[a] - fi[a] >= -w);
      }
      ok = in || !sane;                                                                 // degenerate cell: keep every image
    }
    mask = __ballot_sync(0xffffffffu, ok) & 0x07ffffffu;
    mask |= 1u << 13;                                                                   // the home cell always stays
  }
  LaneImage li;
  const int A = __popc(mask);
  li.q = 32 / A;
  li.jj = lane / A;
  li.live = lane < li.q * A;
  li.img = li.live ? (int)__fns(mask, 0, lane - li.jj * A + 1) : 0;
  return li;
}

// Log-spaced distance bins (4 per octave of d2, from the self-edge threshold 1e-4 up): monotone in d2, so "all
// candidates in bins <= b" is a superset of the cap nearest as soon as those bins hold >= cap candidates.  The count
// pass histograms the in-range candidates per receiver and leaves that bin in the top byte of raw_count; the fill pass
// rejects everything beyond it before it reaches the selection buffer (exactness is unaffected: the selection itself
// still runs on the full (d2, candidate) keys).
constexpr int kBins = 128;
__device__ __forceinline__ int dist_bin(double d2) {
  const int b = (int)(__float_as_uint((float)d2) >> 21) - (int)(0x38D1B717u >> 21);   // 0x38D1B717 = 1e-4f
  return b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
}
constexpr int kCountMask = 0x00FFFFFF;

__device__ __forceinline__ bool cand_pass(double d2, double r2, int remove_self) {
  return (d2 <= r2) && (!remove_self || d2 > 0.0001);   // helpers:432-436
}

template <bool kCull>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
graph_count_kernel(const double* __restrict__ pos, const double* __restrict__ lattice,
                   const int32_t* __restrict__ atom_offset, const int32_t* __restrict__ crystal_of_atom,
                   int N, double r2, int cap, int remove_self, int32_t* __restrict__ raw_count,
                   int32_t* __restrict__ deg, unsigned long long* __restrict__ nimg) {
  __shared__ int s_hist[kWarpsPerBlock][kBins];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + warp;
  if (i >= N) return;
  int* hist = s_hist[warp];
#pragma unroll
  for (int b = 0; b < kBins / 32; ++b) hist[b * 32 + lane] = 0;
  const int g = crystal_of_atom[i];
  const int start = atom_offset[g], n = atom_offset[g + 1] - start;
  const double pix = pos[3 * (size_t)i], piy = pos[3 * (size_t)i + 1], piz = pos[3 * (size_t)i + 2];
  const LaneImage li = active_images<kCull>(lattice + 9 * (size_t)g, pos, start, n, pix, piy, piz, r2, lane);
  double off[3];
  lane_offset(lattice + 9 * (size_t)g, li.img, off);
  __syncwarp();
  int cnt = 0;
  const bool want_hist = cap > 0;
  for (int j0 = 0; j0 < n; j0 += li.q) {
    const int j = j0 + li.jj;
    const bool valid = li.live && j < n;
    double dx, dy, dz;
    const double d2 = lane_d2(pos + 3 * (size_t)(start + (valid ? j : 0)), off, pix, piy, piz, dx, dy, dz);
    if (valid && cand_pass(d2, r2, remove_self)) {
      ++cnt;
      if (want_hist) atomicAdd(&hist[dist_bin(d2)], 1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  int hint = kBins - 1;
  if (want_hist && cnt > cap) {
    // first bin whose cumulative count reaches the cap
    __syncwarp();
    int v[kBins / 32], tot = 0;
#pragma unroll
    for (int b = 0; b < kBins / 32; ++b) { v[b] = hist[lane * (kBins / 32) + b]; tot += v[b]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int run = incl - tot, mine = kBins - 1;
#pragma unroll
    for (int b = kBins / 32 - 1; b >= 0; --b) {
      // scanned from the top so that the lowest qualifying bin of this lane wins
      int upto = run;
      for (int q = 0; q <= b; ++q) upto += v[q];
      if (upto >= cap) mine = lane * (kBins / 32) + b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    hint = mine;
  }
  if (lane == 0) {
    raw_count[i] = cnt | (hint << 24);
    deg[i] = (cap > 0 && cnt > cap) ? cap : cnt;
    // helpers:456-465: _max_neighbors[_max_neighbors > threshold] = threshold, summed per crystal
    // (for cap <= 0 this is the reference's own quirk: zeros / negative numbers)
    const long long clipped = (cnt > cap) ? (long long)cap : (long long)cnt;
    atomicAdd(nimg + g, (unsigned long long)clipped);   // integer atomics: order independent
  }
}

// Single-CTA exclusive scan (N is at most a few 100k atoms per micro-batch): coalesced tiles of 4096 elements
// (int4 per thread), block scan per tile, running carry.
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int n) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 4096) {
    const int carry = carry_s;           // written before the barrier that ended the previous tile
    const int i0 = base + tid * 4;
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = (i0 + k < n) ? in[i0 + k] : 0;
    const int s = (v[0] + v[1]) + (v[2] + v[3]);
    int incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_tot[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;
      if (lane == 31) carry_s = carry + wi;
    }
    __syncthreads();
    int run = carry + warp_tot[warp] + incl - s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (i0 + k < n) out[i0 + k] = run;
      run += v[k];
    }
    __syncthreads();
  }
  if (tid == 0) out[n] = carry_s;
}

struct SelEntry {
  double d2;
  int c;
};

// Keep the `cap` smallest entries by (d2, c) ascending, preserving buffer (= candidate) order.
__device__ __forceinline__ int select_topk(double* sd2, int* sc, unsigned* smask, int m, int cap, int lane) {
  const int chunks = (m + 31) >> 5;
  for (int t = 0; t < chunks; ++t) {
    const int a = t * 32 + lane;
    bool keep = false;
    if (a < m) {
      const double da = sd2[a];
      const int ca = sc[a];
      int rank = 0;
      for (int b = 0; b < m; ++b) {
        const double db = sd2[b];
        rank += (db < da || (db == da && sc[b] < ca)) ? 1 : 0;
      }
      keep = rank < cap;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) smask[t] = bal;
  }
  __syncwarp();
  int out = 0;
  for (int t = 0; t < chunks; ++t) {
    const int a = t * 32 + lane;
    const unsigned bal = smask[t];
    const bool keep = (bal >> lane) & 1u;
    double da = 0.0;
    int ca = 0;
    if (a < m) {
      da = sd2[a];
      ca = sc[a];
    }
    __syncwarp();
    if (keep) {
      const int p = out + __popc(bal & ((1u << lane) - 1u));
      sd2[p] = da;
      sc[p] = ca;
    }
    out += __popc(bal);
    __syncwarp();
  }
  return out;
}

template <bool kCull>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
graph_fill_kernel(const double* __restrict__ pos, const double* __restrict__ lattice,
                  const int32_t* __restrict__ atom_offset, const int32_t* __restrict__ crystal_of_atom, int N,
                  double r2, int cap, int remove_self, const int32_t* __restrict__ raw_count,
                  const int32_t* __restrict__ row_ptr, long long edge_capacity, int32_t* __restrict__ src,
                  int32_t* __restrict__ dst, int8_t* __restrict__ cell, double* __restrict__ dist,
                  double* __restrict__ dir, long long* __restrict__ ei64, double* __restrict__ cell_offsets,
                  int32_t* __restrict__ overflow_flag) {
  __shared__ double s_d2[kWarpsPerBlock][kSelBuf];
  __shared__ int s_c[kWarpsPerBlock][kSelBuf];
  __shared__ unsigned s_mask[kWarpsPerBlock][kSelBuf / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kWarpsPerBlock + warp;
  if (i >= N) return;
  const int g = crystal_of_atom[i];
  const int start = atom_offset[g], n = atom_offset[g + 1] - start;
  const double pix = pos[3 * (size_t)i], piy = pos[3 * (size_t)i + 1], piz = pos[3 * (size_t)i + 2];
  const LaneImage li = active_images<kCull>(lattice + 9 * (size_t)g, pos, start, n, pix, piy, piz, r2, lane);
  double off[3];
  lane_offset(lattice + 9 * (size_t)g, li.img, off);
  const long long base = row_ptr[i];
  const int rc = raw_count[i];
  const bool select = cap > 0 && (rc & kCountMask) > cap;
  const int bin_max = rc >> 24;          // graph_count: the cap nearest all lie in distance bins <= bin_max

  auto emit = [&](long long e, int c, double d2, double dx, double dy, double dz) {
    if (e >= edge_capacity) {
      if (overflow_flag) *overflow_flag = 1;
      return;
    }
    const int j = c / 27, k = c - j * 27;
    src[e] = start + j;
    dst[e] = i;
    cell[e] = (int8_t)k;
    dist[e] = sqrt(d2);                     // helpers:544 (IEEE correctly rounded)
    dir[3 * e] = dx;
    dir[3 * e + 1] = dy;
    dir[3 * e + 2] = dz;
    if (ei64) {
      ei64[e] = start + j;
      ei64[edge_capacity + e] = i;
    }
    if (cell_offsets) {
      int c0, c1, c2;
      cell_of(k, c0, c1, c2);
      cell_offsets[3 * e] = -(double)c0;    // helpers:549 returns -unit_cell
      cell_offsets[3 * e + 1] = -(double)c1;
      cell_offsets[3 * e + 2] = -(double)c2;
    }
  };

  if (!select) {
    int written = 0;
    for (int j0 = 0; j0 < n; j0 += li.q) {
      const int j = j0 + li.jj;
      const bool valid = li.live && j < n;
      double dx, dy, dz;
      const double d2 = lane_d2(pos + 3 * (size_t)(start + (valid ? j : 0)), off, pix, piy, piz, dx, dy, dz);
      const bool ok = valid && cand_pass(d2, r2, remove_self);
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) emit(base + written + __popc(bal & ((1u << lane) - 1u)), 27 * j + li.img, d2, dx, dy, dz);
      written += __popc(bal);
    }
    return;
  }

  double* sd2 = s_d2[warp];
  int* sc = s_c[warp];
  int m = 0;
  // Once the buffer has been cut down to the `cap` nearest seen so far, anything not strictly before the
  // current cap-th entry in (d2, candidate) order can never be selected: reject it before it is buffered.
  // In dense cells (sampler start: ~1 A cells, ~1000 in-range candidates per atom) this keeps the number of
  // O(m^2) selection passes at O(log) instead of one per 224 candidates.
  double thr_d2 = INFINITY;
  int thr_c = 0x7fffffff;
  // compact as soon as cap + 64 candidates are buffered: the selection pass is O(m^2 / 32) per lane
  const int sel_limit = min(kSelBuf, ((cap + 31) / 32) * 32 + 64);
  for (int j0 = 0; j0 < n; j0 += li.q) {
    if (m + 32 > sel_limit) {
      m = select_topk(sd2, sc, s_mask[warp], m, cap, lane);
      // threshold = the largest kept key
      double kd = -INFINITY;
      int kc = -1;
      for (int a = lane; a < m; a += 32)
        if (sd2[a] > kd || (sd2[a] == kd && sc[a] > kc)) { kd = sd2[a]; kc = sc[a]; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, kd, o);
        const int oc = __shfl_xor_sync(0xffffffffu, kc, o);
        if (od > kd || (od == kd && oc > kc)) { kd = od; kc = oc; }
      }
      thr_d2 = kd;
      thr_c = kc;
    }
    const int j = j0 + li.jj;
    const bool valid = li.live && j < n;
    const int c = 27 * j + li.img;
    double dx, dy, dz;
    const double d2 = lane_d2(pos + 3 * (size_t)(start + (valid ? j : 0)), off, pix, piy, piz, dx, dy, dz);
    const bool ok = valid && cand_pass(d2, r2, remove_self) && dist_bin(d2) <= bin_max &&
                    (d2 < thr_d2 || (d2 == thr_d2 && c < thr_c));
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int p = m + __popc(bal & ((1u << lane) - 1u));
      sd2[p] = d2;
      sc[p] = c;
    }
    m += __popc(bal);
    __syncwarp();
  }
  if (m > cap) m = select_topk(sd2, sc, s_mask[warp], m, cap, lane);
  for (int a = lane; a < m; a += 32) {
    const int c = sc[a];
    const int j = c / 27, k = c - 27 * j;
    // the emitting lane is not the image's lane: rebuild the image offset
    double offk[3], dx, dy, dz;
    lane_offset(lattice + 9 * (size_t)g, k, offk);
    const double d2 = lane_d2(pos + 3 * (size_t)(start + j), offk, pix, piy, piz, dx, dy, dz);
    emit(base + a, c, d2, dx, dy, dz);
  }
}

}  // namespace

extern "C" int arreau_graph_count(const double* pos, const double* lattice, const int32_t* atom_offset,
                                  const int32_t* crystal_of_atom, int32_t N, int32_t G, double radius_sq,
                                  int32_t cap, int32_t remove_self_edges, int32_t* raw_count, int32_t* deg,
                                  int64_t* num_neighbors_image, void* stream) {
  if (!pos || !lattice || !atom_offset || !crystal_of_atom || !raw_count || !deg || !num_neighbors_image)
    return ARREAU_ERR_NULL;
  if (N < 0 || G < 0) return ARREAU_ERR_BAD_SHAPE;
  if (cap > kMaxCap) return ARREAU_ERR_UNSUPPORTED;
  cudaStream_t s = (cudaStream_t)stream;
  if (G > 0) {
    cudaError_t e = cudaMemsetAsync(num_neighbors_image, 0, sizeof(int64_t) * (size_t)G, s);
    if (e != cudaSuccess) return (int)e;
  }
  if (N == 0) return ARREAU_OK;
  const int blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if ((long long)N >= (long long)kCullMinAtoms * G)      // large cells on average: the image-culling walk
    graph_count_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, s>>>(pos, lattice, atom_offset, crystal_of_atom, N, radius_sq,
                                                                    cap, remove_self_edges, raw_count, deg,
                                                                    (unsigned long long*)num_neighbors_image);
  else
    graph_count_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, s>>>(pos, lattice, atom_offset, crystal_of_atom, N, radius_sq,
                                                                     cap, remove_self_edges, raw_count, deg,
                                                                     (unsigned long long*)num_neighbors_image);
  CUDA_LAUNCH_CHECK();
  return ARREAU_OK;
}

extern "C" int arreau_graph_scan(const int32_t* deg, int32_t* row_ptr, int32_t n, void* stream) {
  if (!row_ptr || (n > 0 && !deg)) return ARREAU_ERR_NULL;
  if (n < 0) return ARREAU_ERR_BAD_SHAPE;
  scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(deg, row_ptr, n);
  CUDA_LAUNCH_CHECK();
  return ARREAU_OK;
}

extern "C" int arreau_graph_fill(const double* pos, const double* lattice, const int32_t* atom_offset,
                                 const int32_t* crystal_of_atom, int32_t N, int32_t G, double radius_sq,
                                 int32_t cap, int32_t remove_self_edges, const int32_t* raw_count,
                                 const int32_t* row_ptr, int64_t edge_capacity, int32_t* src, int32_t* dst,
                                 int8_t* cell, double* dist, double* dir, int64_t* edge_index_i64,
                                 double* cell_offsets, int32_t* overflow_flag, void* stream) {
  if (N == 0) return ARREAU_OK;
  if (!pos || !lattice || !atom_offset || !crystal_of_atom || !raw_count || !row_ptr) return ARREAU_ERR_NULL;
  if (edge_capacity > 0 && (!src || !dst || !cell || !dist || !dir)) return ARREAU_ERR_NULL;
  if (N < 0 || G < 0 || edge_capacity < 0) return ARREAU_ERR_BAD_SHAPE;
  if (cap > kMaxCap) return ARREAU_ERR_UNSUPPORTED;
  const int blocks = (N + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if ((long long)N >= (long long)kCullMinAtoms * G)
    graph_fill_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        pos, lattice, atom_offset, crystal_of_atom, N, radius_sq, cap, remove_self_edges, raw_count, row_ptr,
        (long long)edge_capacity, src, dst, cell, dist, dir, (long long*)edge_index_i64, cell_offsets, overflow_flag);
  else
    graph_fill_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        pos, lattice, atom_offset, crystal_of_atom, N, radius_sq, cap, remove_self_edges, raw_count, row_ptr,
        (long long)edge_capacity, src, dst, cell, dist, dir, (long long*)edge_index_i64, cell_offsets, overflow_flag);
  CUDA_LAUNCH_CHECK();
  return ARREAU_OK;
}
