// Host replay of the n_fft = 512 kernels (TEST SUPPORT, not a product path and not a fallback:
// nothing in the package loads this library; tests/ builds and calls it).
//
// It compiles csrc/aip_tiles.cuh with g++ and runs the kernels' phase functions thread by thread
// -- for tile: for phase: for tid in 0..255 -- i.e. the same indexing and arithmetic the GPU runs,
// so that the lane/warp -> (frame, job) maps, the exchange-buffer layout, the split pass and the
// overlap-add can be checked against the oracle in the CPU-only test tier.
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "aip_tiles.cuh"

using namespace aip;

extern "C" {

int emul_stft512_fwd(const float* wave, int B, int L, long long pitch, int hop, int center, int win_length,
                     const float* window, const int* gap_samples, const int* zero_frames,
                     const int* mask_frames, int mask_in_gap_is_one, int mag_kind, float eps, float power,
                     int T_out, float* spec, float* mag, float* phase, float* mask, int vec_ok) {
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.wave = wave; P.wave_pitch = pitch; P.B = B; P.L = L;
  P.hop = hop; P.pad = center ? 256 : 0; P.reflect = center == 2;
  P.T = 1 + (L + 2 * P.pad - 512) / hop;
  if (T_out > P.T || (hop & 1)) return -1;
  P.T_out = T_out;
  P.window = window;
  P.gap_samples = gap_samples; P.zero_frames = zero_frames; P.mask_frames = mask_frames;
  P.mask_in_gap_is_one = mask_in_gap_is_one;
  P.mag_kind = mag_kind; P.eps = eps; P.power = power;
  P.spec = reinterpret_cast<float2*>(spec); P.mag = mag; P.phase = phase; P.mask = mask;
  P.tiles_per_clip = (T_out + kFR - 1) / kFR;
  P.n_tiles = B * P.tiles_per_clip;
  P.tile_floats = (fwd_tile_len(hop) + 31) & ~31;
  P.n_tile_bufs = 1;
  P.zero_groups = win_zero_groups(win_length);
  P.vec_ok = vec_ok && ((hop & 3) == 0) && ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
  std::vector<float> tile(P.tile_floats);
  std::vector<float2> exch(kExch);
  std::vector<LaneConst> lc(kThreads);
  std::vector<PairTw> pw(kThreads);
  std::vector<float2> tw_s(kTwTable);
  twiddle_table_fill(tw_s.data(), 0, 1);
  for (int tid = 0; tid < kThreads; ++tid) {
    lane_const_init(lc[tid], tw_s.data(), tid & 15);
    pair_tw_init(pw[tid], tid >> 5);
  }
  NoRelease rel;
  alignas(16) float win_s[kWinTable];
  window_table_fill(win_s, window, 0.5f, 0, 1);
  const int mode = fwd_mode_of(P);
  TileCursor c = tile_cursor(0, P.tiles_per_clip);
  for (int tix = 0; tix < P.n_tiles; ++tix) {
    const FwdTilePlan q = fwd_tile_plan(P, c);
    if (q.n_bulk > 0) memcpy(tile.data() + q.v_lo, q.src + q.g0 + q.v_lo, (size_t)q.n_bulk * 4);   // the TMA bulk copy
    if (fwd_needs_fixup(q)) for (int tid = 0; tid < kThreads; ++tid) fwd_fixup(q, tid, tile.data());
    for (int tid = 0; tid < kThreads; ++tid) {
      if (P.zero_groups == 2) fwd_phase1<2>(P, tid, tile.data(), exch.data(), win_s, lc[tid]);
      else fwd_phase1<0>(P, tid, tile.data(), exch.data(), win_s, lc[tid]);
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      switch (mode) {
#define AIP_CASE(M) case (M): fwd_phase2<(M)>(P, tid, c, exch.data(), pw[tid], rel); break;
        AIP_CASE(FWD_MAG_ABS) AIP_CASE(FWD_MAG_LOG10) AIP_CASE(FWD_SPEC) AIP_CASE(MAG_LOG10_EPS | FWD_MASK)
        AIP_CASE(MAG_LOG1P_POW) AIP_CASE(MAG_LOG1P_POW | FWD_PHASE | FWD_MASK) AIP_CASE(FWD_SPEC | FWD_PHASE | FWD_MASK)
        AIP_CASE(MAG_LOG10_EPS | FWD_ZERO) AIP_CASE(MAG_ABS | FWD_PHASE) AIP_CASE(MAG_LOG1P_POW | FWD_PHASE) AIP_CASE(FWD_SPEC | FWD_PHASE)
        AIP_CASE(MAG_POW) AIP_CASE(FWD_SPEC | MAG_ABS) AIP_CASE(FWD_SPEC | MAG_LOG10_EPS) AIP_CASE(FWD_SPEC | MAG_LOG1P_POW)
        AIP_CASE(MAG_LOG10_EPS | FWD_PHASE)
#undef AIP_CASE
        default: fwd_phase2<FWD_FULL>(P, tid, c, exch.data(), pw[tid], rel); break;
      }
    }
    tile_advance(c, P.tiles_per_clip);
  }
  return 0;
}

// aip_stft_gap_variants_f32: the copy pass (plain loops) and the variant tiles, replayed thread by thread
int emul_stft512_variants(const float* wave, int N, int L, long long pitch, int hop, int center, int win_length,
                          const float* window, int G, const int* gap_samples, int gap_len_max, int mag_kind, float eps,
                          int T_out, const float* clean_mag, float* mag, int vec_ok) {
  FwdParams P;
  memset(&P, 0, sizeof(P));
  P.wave = wave; P.wave_pitch = pitch; P.B = N * G; P.L = L;
  P.hop = hop; P.pad = center ? 256 : 0; P.reflect = center == 2;
  P.T = 1 + (L + 2 * P.pad - 512) / hop;
  if (T_out > P.T || (hop & 1)) return -1;
  P.T_out = T_out;
  P.window = window;
  P.gap_samples = gap_samples;
  P.mag_kind = mag_kind; P.eps = eps; P.power = 1.0f; P.mag = mag;
  P.var_div = G;
  P.tiles_per_clip = var_tiles(gap_len_max, hop, T_out);
  P.n_tiles = P.B * P.tiles_per_clip;
  P.tile_floats = (fwd_tile_len(hop) + 31) & ~31;
  P.n_tile_bufs = 1;
  P.zero_groups = win_zero_groups(win_length);
  P.vec_ok = vec_ok && ((hop & 3) == 0) && ((pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(wave) & 15) == 0);
  const long long FT = (long long)kBins * T_out;
  for (int v = 0; v < P.B; ++v) memcpy(mag + v * FT, clean_mag + (v / G) * FT, (size_t)FT * sizeof(float));
  std::vector<float> tile(P.tile_floats);
  std::vector<float2> exch(kExch);
  std::vector<LaneConst> lc(kThreads);
  std::vector<PairTw> pw(kThreads);
  std::vector<float2> tw_s(kTwTable);
  twiddle_table_fill(tw_s.data(), 0, 1);
  for (int tid = 0; tid < kThreads; ++tid) {
    lane_const_init(lc[tid], tw_s.data(), tid & 15);
    pair_tw_init(pw[tid], tid >> 5);
  }
  NoRelease rel;
  alignas(16) float win_s[kWinTable];
  window_table_fill(win_s, window, 0.5f, 0, 1);
  TileCursor c = tile_cursor(0, P.tiles_per_clip);
  for (int tix = 0; tix < P.n_tiles; ++tix) {
    const int gs = gap_samples[2 * c.b], ge = gap_samples[2 * c.b + 1];
    const FwdTilePlan q = fwd_tile_plan_var(P, c, gs, ge, var_frame_base(P, gs), c.b / G);      // = variant_meta_kernel
    if (q.n_bulk > 0) memcpy(tile.data() + q.v_lo, q.src + q.g0 + q.v_lo, (size_t)q.n_bulk * 4);   // the TMA bulk copy
    const bool own = !fwd_needs_edge_fixup(q);      // barrier-free path: each warp zeroes the gap for its own frames only
    if (!own && fwd_needs_fixup(q)) for (int tid = 0; tid < kThreads; ++tid) fwd_fixup(q, tid, tile.data());
    std::vector<float> mine;
    for (int w = 0; w < kThreads / 32; ++w) {
      // replayed warp by warp on a PRIVATE copy of the staged tile: a warp must not depend on another warp's zeroing
      float* t = tile.data();
      if (own) {
        mine = tile;
        t = mine.data();
        for (int tid = 32 * w; tid < 32 * w + 32; ++tid) fwd_gap_zero_own(q, P.hop, tid, t);
      }
      for (int tid = 32 * w; tid < 32 * w + 32; ++tid) {
        if (P.zero_groups == 2) fwd_phase1<2>(P, tid, t, exch.data(), win_s, lc[tid]);
        else fwd_phase1<0>(P, tid, t, exch.data(), win_s, lc[tid]);
      }
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      switch (mag_kind) {
        case MAG_ABS: fwd_phase2<FWD_MAG_ABS | FWD_VARIANT>(P, tid, c, exch.data(), pw[tid], rel, var_frame_base(P, gs)); break;
        case MAG_LOG10_EPS: fwd_phase2<FWD_MAG_LOG10 | FWD_VARIANT>(P, tid, c, exch.data(), pw[tid], rel, var_frame_base(P, gs)); break;
        case MAG_LOG1P_POW: fwd_phase2<MAG_LOG1P_POW | FWD_VARIANT>(P, tid, c, exch.data(), pw[tid], rel, var_frame_base(P, gs)); break;
        default: return -2;
      }
    }
    tile_advance(c, P.tiles_per_clip);
  }
  return 0;
}

// the epilogue's scalar helpers, for the accuracy tests
void emul_fast_math(int n, const float* y, const float* x, float* out_atan2, float* out_log1p) {
  for (int i = 0; i < n; ++i) {
    out_atan2[i] = fast_atan2(y[i], x[i]);
    out_log1p[i] = fast_log1p(fabsf(x[i]));
  }
}

void emul_fast_expm1(int n, const float* x, float* out) {
  for (int i = 0; i < n; ++i) out[i] = fast_expm1(x[i]);
}

int emul_istft512(const float* spec, const float* mag, const float* phase, int mag_domain,
                  const int* db_flags, int B, int T, int length, int hop, int center, int win_length,
                  const float* window, const float* inv_wss, float* out, long long out_pitch,
                  const float* blend_in, const float* blend_mask, float* peaks) {
  InvParams P;
  memset(&P, 0, sizeof(P));
  P.blend_in = blend_in; P.blend_mask = blend_mask; P.peaks = peaks;
  P.spec = reinterpret_cast<const float2*>(spec); P.mag = mag; P.phase = phase; P.mag_domain = mag_domain;
  P.db_flags = db_flags; P.B = B; P.T = T;
  P.hop = hop; P.pad = center ? 256 : 0;
  if (hop & 1) return -1;
  P.g = inv_geom(hop, P.pad);
  if (P.g.FO < 4) return -2;
  int out_len = length > 0 ? length : 512 + hop * (T - 1) - 2 * P.pad;
  P.n_frames = T;
  if (length > 0) {
    const int nf = (length + 2 * P.pad + hop - 1) / hop;
    P.n_frames = nf < T ? nf : T;
  }
  P.out_len = out_len;
  P.window = window; P.inv_wss = inv_wss; P.out = out; P.out_pitch = out_pitch;
  P.vec_ok = ((out_pitch & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0) &&
             ((reinterpret_cast<uintptr_t>(inv_wss) & 7) == 0);
  const long long span = (long long)P.g.FO * hop;
  P.tiles_per_clip = (int)((out_len + span - 1) / span);
  P.n_tiles = B * P.tiles_per_clip;
  P.hop_magic = (unsigned)((0x100000000ULL + (unsigned)hop - 1) / (unsigned)hop);
  P.col_magic = (unsigned)((0x100000000ULL + (unsigned)(hop / 2) - 1) / (unsigned)(hop / 2));
  P.ola_terms = (kNfft + hop - 1) / hop;
  P.ola_dq = (2 * kThreads) / hop;
  P.ola_dr = (2 * kThreads) % hop;
  {
    int f_ref = P.ola_terms - 1;
    const int need = (P.pad + hop - 1) / hop;
    if (f_ref < need) f_ref = need;
    const long long s_ref = (long long)f_ref * hop - P.pad;
    P.wss_ref = (f_ref <= P.n_frames - 1 && s_ref >= 0 && s_ref + hop <= out_len) ? (int)s_ref : -1;
  }
  P.ola_fast = inv_ola_fast_kind(hop, P.pad, win_length);
  std::vector<float> wtab;
  if (P.wss_ref >= 0 && hop <= kMaxWtab) wtab.assign(inv_wss + P.wss_ref, inv_wss + P.wss_ref + hop);
  std::vector<float2> exch(kExch);
  std::vector<LaneConst> lc(kThreads);
  std::vector<PairTw> pw(kThreads);
  std::vector<float2> tw_s(kTwTable);
  twiddle_table_fill(tw_s.data(), 0, 1);
  for (int tid = 0; tid < kThreads; ++tid) {
    lane_const_init(lc[tid], tw_s.data(), tid & 15);
    pair_tw_init(pw[tid], tid >> 5);
  }
  NoRelease rel;
  alignas(16) float win_s[kWinTable];
  window_table_fill(win_s, window, 1.0f / 512.0f, 0, 1);
  TileCursor c = tile_cursor(0, P.tiles_per_clip);
  for (int tix = 0; tix < P.n_tiles; ++tix) {
    for (int tid = 0; tid < kThreads; ++tid) {
      switch (inv_mode_of(P)) {
#define AIP_CASE(M) case (M): inv_phase0<(M)>(P, tid, c, exch.data(), pw[tid], rel); break;
        AIP_CASE(INV_SPEC) AIP_CASE(1) AIP_CASE(2) AIP_CASE(3) AIP_CASE(4) AIP_CASE(5) AIP_CASE(6) AIP_CASE(INV_BLEND) AIP_CASE(INV_BLEND_LIN) AIP_CASE(INV_BLEND_EXPM1)
#undef AIP_CASE
      }
    }
    const float* wt = wtab.empty() ? nullptr : wtab.data();
    for (int tid = 0; tid < kThreads; ++tid) {
      if (P.ola_fast == 1) inv_phase1<1>(P, tid, c, exch.data(), win_s, lc[tid]);
      else if (P.ola_fast == 2) inv_phase1<2>(P, tid, c, exch.data(), win_s, lc[tid]);
      else inv_phase1<0>(P, tid, c, exch.data(), win_s, lc[tid]);
    }
    for (int tid = 0; tid < kThreads; ++tid) {
      if (P.ola_fast == 1) inv_phase2<1>(P, tid, c, exch.data(), wt);
      else if (P.ola_fast == 2) inv_phase2<2>(P, tid, c, exch.data(), wt);
      else inv_phase2<0>(P, tid, c, exch.data(), wt);
    }
    tile_advance(c, P.tiles_per_clip);
  }
  return 0;
}

}  // extern "C"
