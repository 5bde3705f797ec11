// timelag_pair.cuh -- the time-lag sweep with TWO FRAMES PER LANE (included by timelag.cu).
//
// Same arithmetic and outputs as timelag_kernel (sums[a][q][f] = sum_e m_e x[f][a][e] x[f-w_q][a][e], q = 0 is the
// squared norm; reference semantics timeSOAP, src/SOAPify/analysis.py:55-60).  timelag_kernel runs at ~96 % of the
// cycles an SM's shared memory has: per 16-byte unit x 32 frames it reads 4 (own frame) + 4 per window
// + 1 (multiplicities) wavefronts.  Here a lane owns the frame PAIR (2p, 2p+1):
//   * the lag-1 partner of the odd frame is the lane's own even frame (a register): window 1 costs 2 instead of 4;
//   * the multiplicities are read once per pair (0.5 instead of 1).
// Even and odd frames live in separate halves of the ring (two boxes of a 3-D tensor map [F/2][2][A*D] per
// 64-frame block), so that consecutive lanes read consecutive slots of a half: with slots 16*odd bytes apart every
// 128-bit load stays conflict free, for the lag partners too (frame 2p - w sits in the half of its parity at slot
// p - ceil(w/2) or p - floor(w/2)).  A block is 64 frames; the ring holds three (previous = lag history up to 64
// frames, current, next in flight), which leaves pieces of ~61 units, i.e. more (atom, piece) items and more
// partial-sum traffic per unit than the 103-unit pieces of timelag_kernel -- the microbenchmark of the consumer
// loops (scripts/micro/lds_pair.cu against lds_block.cu) reads 21.9 against 25.0 clk per unit x 32 frames.
// Applies to: TMA-able rows of whole units (no shifted windows), an even number of frames, cosine-type kinds,
// windows <= 64 with window 1 among them.  Everything else takes timelag_kernel.
#pragma once

namespace soapb {

constexpr int kPrBlk = 64;        // frames per block (32 pairs = 32 lanes)
constexpr int kPrSlots = 3;       // blocks in the ring: previous, current, next
constexpr int kPrRP = 32 * kPrSlots;  // pair slots per parity half
constexpr int kPrWarps = 8;

struct PairPlan {
  int unit_elems, n_units, split, upp, stride, Q, mbuf;
  size_t smem;
};

static size_t pr_smem_bytes(int stride, int upp, int Q, int mbuf) {
  return (size_t)2 * kPrRP * stride + (size_t)mbuf * upp * 16 + (size_t)kPrWarps * Q * kPrBlk * 8 + (size_t)(2 * kPrSlots + 2) * 8 + 64;
}

static bool pr_make_plan(int64_t F, int D, int elem_size, int nw, PairPlan &pl) {
  pl.unit_elems = 16 / elem_size;
  pl.n_units = (D + pl.unit_elems - 1) / pl.unit_elems;
  pl.Q = 1 + nw;
  const size_t budget = (size_t)max_smem_optin() - 1024;
  const int64_t nblk = (F + kPrBlk - 1) / kPrBlk;
  pl.mbuf = nblk >= 1 ? 2 : 4;
  for (int split = 1; split <= pl.n_units; ++split) {
    const int upp = ((pl.n_units + split - 1) / split) | 1;
    if (upp > 127) continue;
    const int stride = 16 * upp;
    const size_t smem = pr_smem_bytes(stride, upp, pl.Q, pl.mbuf);
    if (smem <= budget) {
      pl.split = (pl.n_units + upp - 1) / upp;
      pl.upp = upp;
      pl.stride = stride;
      pl.smem = smem;
      return true;
    }
  }
  return false;
}

// 3-D view of X[F][A*D] as [F/2][2][A*D] 8-byte words: box = one parity x 32 pairs x one piece
inline int make_tensor_map_pairs(CUtensorMap *map, const void *base, int64_t n_pairs, int64_t row_words, int box_words) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return set_error(SOAP_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {(cuuint64_t)row_words, 2, (cuuint64_t)n_pairs};
  cuuint64_t strides[2] = {(cuuint64_t)row_words * 8, (cuuint64_t)row_words * 16};
  cuuint32_t box[3] = {(cuuint32_t)box_words, 1, 32};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SOAP_ECUDA, "cuTensorMapEncodeTiled (pairs) failed with code %d", (int)r);
  return SOAP_OK;
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// U units of a frame pair against the NW lagged frames of each.  acc[0 .. 2Q) belong to the even frame, acc[2Q .. 4Q)
// to the odd one (slots 2q and 2q + 1 of a sum: the float32 variant keeps two chains per sum, the float64 one uses the
// first only).  W1: the first window is 1 -- the odd frame's partner is the even frame itself and ye[0] addresses
// frame 2p - 1.
template <int NW, int U, bool HM, bool W1>
__device__ __forceinline__ void accum_pair_f64(uint32_t m_addr, uint32_t xe, uint32_t xo, const uint32_t (&ye)[NW], const uint32_t (&yo)[NW],
                                               uint32_t off, uint32_t ustep, double (&acc)[4 * (NW + 1)]) {
  constexpr int Q2 = 2 * (NW + 1);
  // Two phases, so that at most 3 + NW loads per unit are live at once (the kernel sits at the 168-register cap of a
  // 10-warp CTA): first the even frame against its partners (and everything the odd frame needs from registers),
  // then the odd frame's partners.  ONE chain per sum (timelag_kernel keeps two): (NW + 1) sums x 2 frames are
  // already independent chains.
  double2 m[U], a[U], b[U], pe[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const uint32_t o = off + j * ustep;
    m[j] = HM ? lds_d2(m_addr + o) : make_double2(1.0, 1.0);
    a[j] = lds_d2(xe + o);
    b[j] = lds_d2(xo + o);
#pragma unroll
    for (int k = 0; k < NW; ++k) pe[k][j] = lds_d2(ye[k] + o);
  }
  double2 bm[U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double am0 = a[j].x * m[j].x, am1 = a[j].y * m[j].y;
    bm[j] = make_double2(b[j].x * m[j].x, b[j].y * m[j].y);
    acc[0] = fma(am1, a[j].y, fma(am0, a[j].x, acc[0]));
    acc[Q2] = fma(bm[j].y, b[j].y, fma(bm[j].x, b[j].x, acc[Q2]));
#pragma unroll
    for (int k = 0; k < NW; ++k) acc[2 * k + 2] = fma(am1, pe[k][j].y, fma(am0, pe[k][j].x, acc[2 * k + 2]));
    if (W1) acc[Q2 + 2] = fma(bm[j].y, a[j].y, fma(bm[j].x, a[j].x, acc[Q2 + 2]));
  }
#pragma unroll
  for (int k = W1 ? 1 : 0; k < NW; ++k) {
    double2 po[U];
#pragma unroll
    for (int j = 0; j < U; ++j) po[j] = lds_d2(yo[k] + off + j * ustep);
#pragma unroll
    for (int j = 0; j < U; ++j) acc[Q2 + 2 * k + 2] = fma(bm[j].y, po[j].y, fma(bm[j].x, po[j].x, acc[Q2 + 2 * k + 2]));
  }
}

// fp32 data: products and short chains in fp32, chain sums promoted to double (as accum_units_f32)
template <int NW, int U, bool HM, bool W1>
__device__ __forceinline__ void accum_pair_f32(uint32_t m_addr, uint32_t xe, uint32_t xo, const uint32_t (&ye)[NW], const uint32_t (&yo)[NW],
                                               uint32_t off, uint32_t ustep, double (&acc)[4 * (NW + 1)]) {
  constexpr int Q2 = 2 * (NW + 1);
  float4 m[U], a[U], b[U], pe[NW][U], po[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const uint32_t o = off + j * ustep;
    m[j] = HM ? lds_f4(m_addr + o) : make_float4(1.f, 1.f, 1.f, 1.f);
    a[j] = lds_f4(xe + o);
    b[j] = lds_f4(xo + o);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      pe[k][j] = lds_f4(ye[k] + o);
      if (!(W1 && k == 0)) po[k][j] = lds_f4(yo[k] + o);
    }
  }
  auto chain = [&](const float4 (&x)[U], const float4 (&y)[U], double &s0, double &s1) {
    float c0 = 0.f, c1 = 0.f;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      c0 = fmaf(x[j].x * m[j].x, y[j].x, c0);
      c1 = fmaf(x[j].y * m[j].y, y[j].y, c1);
      c0 = fmaf(x[j].z * m[j].z, y[j].z, c0);
      c1 = fmaf(x[j].w * m[j].w, y[j].w, c1);
    }
    s0 += (double)c0;
    s1 += (double)c1;
  };
  chain(a, a, acc[0], acc[1]);
  chain(b, b, acc[Q2], acc[Q2 + 1]);
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    chain(a, pe[k], acc[2 * k + 2], acc[2 * k + 3]);
    if (W1 && k == 0)
      chain(b, a, acc[Q2 + 2], acc[Q2 + 3]);
    else
      chain(b, po[k], acc[Q2 + 2 * k + 2], acc[Q2 + 2 * k + 3]);
  }
}

// Warp roles as in timelag_kernel: warps 0..7 consume (lane = frame pair), warp 8 produces (two tensor-map boxes per
// 64-frame block), warp 9 reduces the consumers' partial sums of a block (and the pieces of an atom, in the L2-resident
// per-CTA accumulator).  full[s] / empty[s] per ring block, red_full / red_empty for the partial sums.
template <typename T, int NW, bool HAS_MULT, bool W1, int UB>
__global__ void __launch_bounds__((kPrWarps + 2) * 32, 1)
    timelag_pair_kernel(const T *__restrict__ mult, double *__restrict__ sums, double *__restrict__ cta_acc, long long F, long long A, int D,
                        int split, int upp, int stride, int mbuf, int4 win4, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int UE = 16 / (int)sizeof(T);
  constexpr int Q = NW + 1;
  constexpr int CW = kPrWarps;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};

  unsigned char *ring_e = smem_raw;                                  // even frames: [kPrRP][stride]
  unsigned char *ring_o = ring_e + (size_t)kPrRP * stride;           // odd frames
  unsigned char *msm = ring_o + (size_t)kPrRP * stride;              // [mbuf][upp] units
  double *red = reinterpret_cast<double *>(msm + (size_t)mbuf * upp * 16);  // [CW][Q][64]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + CW * Q * kPrBlk);     // [kPrSlots]
  uint64_t *empty = full + kPrSlots;
  uint64_t *red_full = empty + kPrSlots;
  uint64_t *red_empty = red_full + 1;

  const bool atom_major = cta_acc != nullptr;
  const long long my_atoms = (long long)blockIdx.x < A ? (A - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_items = A * split;
  const int n_my = atom_major ? (int)my_atoms * split
                              : ((long long)blockIdx.x < n_items ? (int)((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0);
  const int nblk = (int)((F + kPrBlk - 1) / kPrBlk);

  if (tid == 0) {
    for (int s = 0; s < kPrSlots; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CW);
    }
    mbar_init(red_full, CW);
    mbar_init(red_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == CW) {
    // ===== producer: block g overwrites block g - 3, last needed (as lag history) by block g - 2 =====================
    int s = 0, js = 0, jph = 0, g = 0;
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, false);
      const int mslot = k & (mbuf - 1);
      const T *mrow = mult + it.e0;
      for (int b = 0; b < nblk; ++b, ++g) {
        if (g >= kPrSlots - 1) {
          mbar_wait(&empty[js], (uint32_t)jph);
          if (++js == kPrSlots) js = 0, jph ^= 1;
        }
        if (lane == 0) {
          const uint32_t mbytes = (HAS_MULT && b == 0) ? (uint32_t)it.punits * 16u : 0u;
          mbar_expect_tx(&full[s], 2u * 32u * (uint32_t)stride + mbytes);
          tma_load_3d(ring_e + (size_t)s * 32 * stride, &tmap, (int)it.w0, 0, b * 32, &full[s]);
          tma_load_3d(ring_o + (size_t)s * 32 * stride, &tmap, (int)it.w0, 1, b * 32, &full[s]);
          if (mbytes) bulk_g2s(msm + (size_t)mslot * upp * 16, mrow, mbytes, &full[s]);
        }
        if (++s == kPrSlots) s = 0;
      }
    }
  } else if (warp == CW + 1) {
    // ===== reducer: 64 frames per block, two per lane ================================================================
    int g = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, false);
      double *srow = atom_major ? sums + (size_t)it.a * Q * (size_t)F : sums + ((size_t)it.a * split + it.p) * Q * (size_t)F;
      double *acc_row = cta_acc + (size_t)blockIdx.x * Q * (size_t)nblk * kPrBlk;
      const bool first_piece = !atom_major || it.p == 0, last_piece = !atom_major || it.p == split - 1;
      for (int b = 0; b < nblk; ++b, ++g) {
        double prev[2][Q];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int q = 0; q < Q; ++q)
            prev[h][q] = first_piece ? 0.0 : ld_acc(acc_row + (size_t)q * nblk * kPrBlk + (size_t)b * kPrBlk + 32 * h + lane, pol);
        mbar_wait(red_full, (uint32_t)(g & 1));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int fb = 32 * h + lane;  // frame within the block
          const long long fl = (long long)b * kPrBlk + fb;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            double sum = 0.0;
#pragma unroll
            for (int w = 0; w < CW; ++w) sum += red[(w * Q + q) * kPrBlk + fb];
            sum = first_piece ? sum : prev[h][q] + sum;  // ((s_0 + s_1) + s_2) + ...: fixed order
            if (last_piece) {
              if (fl < F) srow[(size_t)q * F + fl] = sum;
            } else {
              st_acc(acc_row + (size_t)q * nblk * kPrBlk + fl, sum, pol);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(red_empty);
      }
    }
  } else {
    // ===== consumers: lane <-> frame pair ============================================================================
    const uint32_t re_addr = smem_u32(ring_e), ro_addr = smem_u32(ring_o), msm_addr = smem_u32(msm);
    const uint32_t ustep = 16u * CW, off0 = 16u * warp;
    // frame 2p - w (partner of the even frame): w even -> even half, slot p - w/2; w odd -> odd half, slot p - (w+1)/2
    // frame 2p + 1 - w (partner of the odd frame): w even -> odd half, slot p - w/2; w odd -> even half, slot p - (w-1)/2
    int de[NW], dodd[NW];
    bool e_in_odd[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      const int w = win[k];
      e_in_odd[k] = (w & 1) != 0;
      de[k] = (w & 1) ? (w + 1) / 2 : w / 2;
      dodd[k] = (w & 1) ? (w - 1) / 2 : w / 2;
    }
    int s = 0, ph = 0, g = 0;
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, false);
      const int my_units = it.punits > warp ? (it.punits - warp + CW - 1) / CW : 0;
      const uint32_t m_addr = msm_addr + (uint32_t)(k & (mbuf - 1)) * upp * 16;
      for (int b = 0; b < nblk; ++b, ++g) {
        const int ps = s * 32 + lane;  // pair slot of this lane in both halves
        const uint32_t xe = re_addr + (uint32_t)ps * stride, xo = ro_addr + (uint32_t)ps * stride;
        uint32_t ye[NW], yo[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          int a = ps - de[q], c = ps - dodd[q];
          if (a < 0) a += kPrRP;
          if (c < 0) c += kPrRP;
          ye[q] = (e_in_odd[q] ? ro_addr : re_addr) + (uint32_t)a * stride;  // frames f < w read stale data: result unused
          yo[q] = (e_in_odd[q] ? re_addr : ro_addr) + (uint32_t)c * stride;
        }
        double acc[4 * Q];
#pragma unroll
        for (int q = 0; q < 4 * Q; ++q) acc[q] = 0.0;
        mbar_wait(&full[s], (uint32_t)ph);
        int n = 0;
        for (; n + UB <= my_units; n += UB) {
          if constexpr (sizeof(T) == 8)
            accum_pair_f64<NW, UB, HAS_MULT, W1>(m_addr, xe, xo, ye, yo, off0 + n * ustep, ustep, acc);
          else
            accum_pair_f32<NW, UB, HAS_MULT, W1>(m_addr, xe, xo, ye, yo, off0 + n * ustep, ustep, acc);
        }
        for (; n < my_units; ++n) {
          if constexpr (sizeof(T) == 8)
            accum_pair_f64<NW, 1, HAS_MULT, W1>(m_addr, xe, xo, ye, yo, off0 + n * ustep, ustep, acc);
          else
            accum_pair_f32<NW, 1, HAS_MULT, W1>(m_addr, xe, xo, ye, yo, off0 + n * ustep, ustep, acc);
        }
        if (g >= 1) mbar_wait(red_empty, (uint32_t)((g - 1) & 1));
        // frame 2 * lane (even) and 2 * lane + 1 (odd) of the block: adjacent doubles, one 128-bit store per sum
        double *r = red + (size_t)warp * Q * kPrBlk;
#pragma unroll
        for (int q = 0; q < Q; ++q)
          *reinterpret_cast<double2 *>(r + q * kPrBlk + 2 * lane) =
              make_double2(acc[2 * q] + acc[2 * q + 1], acc[2 * Q + 2 * q] + acc[2 * Q + 2 * q + 1]);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(red_full);
          mbar_arrive(&empty[s]);
        }
        if (++s == kPrSlots) s = 0, ph ^= 1;
      }
    }
  }
}

}  // namespace soapb
