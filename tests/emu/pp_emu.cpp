// extern "C" launchers of the csrc/pp_ops.cuh kernels under the host emulation (tests/emu/cuda_emu.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_emu.h"
#include "../../video-subtitle-remover_b200/csrc/pp_ops.cuh"
using namespace vsr;
typedef const __half* ch;
typedef __half* mh;
#define L(n, threads, lock, call) emu_launch(n, threads, lock, [&] { call; })

extern "C" {
void emu_frames_to_half(const uint8_t* bgr, size_t px, mh out) { L(dim3(emu_blocks(px)), 256, false, pp_frames_to_half_kernel(bgr, px, out)); }
void emu_instnorm(ch x, int N, size_t px, int cp, int relu, float* mean, float* rstd, mh out) {
  L(dim3(cp / 8, N), 256, true, pp_instnorm_stats_kernel(x, px, cp, mean, rstd));
  const size_t t8 = (size_t)N * px * (cp / 8);
  L(dim3(emu_blocks(t8)), 256, false, pp_instnorm_apply_kernel(x, px, cp, mean, rstd, relu, out, t8));
}
void emu_context_split(ch x, size_t px, mh net, int pn, mh inp, int pi) { L(dim3(emu_blocks(px * 32)), 256, false, pp_context_split_kernel(x, px, net, pn, inp, pi)); }
void emu_corr_pool(ch in, size_t rows, int h2, int w2, int pin, mh out, int pout) {
  L(dim3(emu_blocks(rows * (h2 / 2) * (w2 / 2))), 256, false, pp_corr_pool_kernel(in, rows, h2, w2, pin, out, pout));
}
void emu_corr_lookup(ch l0, ch l1, ch l2, ch l3, const int* hs, const int* ws, const int* ps, const float* flow, int h, int w, size_t px, mh out, int pitch) {
  CorrLevels lv;
  ch p[4] = {l0, l1, l2, l3};
  for (int i = 0; i < 4; ++i) { lv.ptr[i] = p[i]; lv.h[i] = hs[i]; lv.w[i] = ws[i]; lv.pitch[i] = ps[i]; }
  L(dim3(emu_blocks(px * 324)), 256, false, pp_corr_lookup_kernel(lv, flow, h, w, px, out, pitch));
}
void emu_gru_rh(ch r, int pr, ch hs, int ph, mh out, int po, size_t px) { L(dim3(emu_blocks(px * 16)), 256, false, pp_gru_rh_kernel(r, pr, hs, ph, out, po, px)); }
void emu_gru_update(ch z, int pz, ch q, int pq, mh hio, int ph, size_t px) { L(dim3(emu_blocks(px * 16)), 256, false, pp_gru_update_kernel(z, pz, q, pq, hio, ph, px)); }
void emu_flow_update(float* f32, ch delta, int pd, mh f16, mh a, mh b, int pab, int coff, size_t px, int add) {
  L(dim3(emu_blocks(px)), 256, false, pp_flow_update_kernel(f32, delta, pd, f16, a, b, pab, coff, px, add));
}
void emu_convex_upsample(const float* f32, ch mask, int pm, int N, int h, int w, float* out) {
  L(dim3(emu_blocks((size_t)N * h * w * 64)), 256, false, pp_convex_upsample_kernel(f32, mask, pm, N, h, w, out));
}
void emu_img_prop_step(ch prev, ch cur, const float* fp, const float* fc, int H, int W, mh out) {
  L(dim3((W + 255) / 256, H), 256, false, pp_img_prop_step_kernel(prev, cur, fp, fc, H, W, out));
}
void emu_state_init(ch frames, const uint8_t* mask, size_t plane, size_t px, mh out) { L(dim3(emu_blocks(px)), 256, false, pp_state_init_kernel(frames, mask, plane, px, out)); }
void emu_state_compose(ch frames, const uint8_t* mask, ch prop, size_t plane, size_t px, mh out) {
  L(dim3(emu_blocks(px)), 256, false, pp_state_compose_kernel(frames, mask, prop, plane, px, out));
}
void emu_rfc_input(const float* flow, const uint8_t* mask, int N, size_t plane, int rev, mh out) {
  L(dim3(emu_blocks(plane * N)), 256, false, pp_rfc_input_kernel(flow, mask, N, plane, rev, out));
}
void emu_pad_replicate(ch in, int T, int H, int W, int cp, mh out, int OH, int OW, int top, int left) {
  L(dim3(emu_blocks((size_t)T * OH * OW * (cp / 8))), 256, false, pp_pad_replicate_kernel(in, T, H, W, cp, out, OH, OW, top, left));
}
void emu_leaky(mh x, size_t n8, float slope) { L(dim3(emu_blocks(n8)), 256, false, pp_leaky_relu_kernel(x, n8, slope)); }
void emu_temporal_taps(ch in, int T, size_t px, int cpi, mh out, int cpo) {
  L(dim3(emu_blocks((size_t)T * px * 3 * (cpi / 8))), 256, false, pp_temporal_taps_kernel(in, T, px, cpi, out, cpo));
}
void emu_deform_cols(ch xa, int pa, int Ca, ch xb, int pb, int C, int G, ch om, int pom, float maxres, const float* flow, int H, int W, size_t px, mh cols, int pc) {
  L(dim3(emu_blocks(px * G * 9)), 256, false, pp_deform_cols_kernel(xa, pa, Ca, xb, pb, C, G, om, pom, maxres, flow, H, W, px, cols, pc));
}
void emu_rfc_combine(ch pred, int pp, const float* flow, const uint8_t* mask, int N, size_t plane, int rev, float* out) {
  L(dim3(emu_blocks(plane * N)), 256, false, pp_rfc_combine_kernel(pred, pp, flow, mask, N, plane, rev, out));
}
void emu_gen_input(ch state, const uint8_t* mask, const int* ids, int n, size_t plane, mh out) {
  L(dim3(emu_blocks(plane * n)), 256, false, pp_gen_input_kernel(state, mask, ids, n, plane, out));
}
void emu_flow_down4(const float* flow, const int* ids, int n, int H, int W, float* out) {
  L(dim3(emu_blocks((size_t)n * (H / 4) * (W / 4))), 256, false, pp_flow_down4_kernel(flow, ids, n, H, W, out));
}
void emu_prop_masks(ch gin, int n, int H, int W, mh out) { L(dim3(emu_blocks((size_t)n * (H / 4) * (W / 4))), 256, false, pp_prop_masks_kernel(gin, n, H, W, out)); }
void emu_featprop_cond(ch prop, ch cur, int C, const float* fp, const float* fc, ch masks, int H, int W, mh cond, int pitch) {
  L(dim3(emu_blocks((size_t)H * W * (C / 8))), 256, false, pp_featprop_cond_kernel(prop, cur, C, fp, fc, masks, H, W, cond, pitch));
}
void emu_write_extra(ch src, mh dst, int pitch, int coff, int nch, size_t px) { L(dim3(emu_blocks(px)), 256, false, pp_write_extra_kernel(src, dst, pitch, coff, nch, px)); }
void emu_unfold7s3(ch in, int n, int h, int w, int C, mh out, int pitch, int gelu) {
  const int fh = (h + 6 - 7) / 3 + 1, fw = (w + 6 - 7) / 3 + 1;
  L(dim3(emu_blocks((size_t)n * fh * fw * 49 * (C / 8))), 256, false, pp_unfold7s3_kernel(in, n, h, w, C, fh, fw, out, pitch, gelu));
}
void emu_fold7s3(ch tok, int n, int h, int w, int C, int pitch, int norm, mh out) {
  const int fh = (h + 6 - 7) / 3 + 1, fw = (w + 6 - 7) / 3 + 1;
  L(dim3(emu_blocks((size_t)n * h * w * (C / 8))), 256, false, pp_fold7s3_kernel(tok, n, fh, fw, pitch, C, h, w, norm, out));
}
void emu_layernorm(ch x, size_t tokens, int C, const float* g, const float* b, mh out) {
  L(dim3((unsigned)((tokens + 7) / 8)), 256, true, pp_layernorm_kernel(x, tokens, C, g, b, out));
}
void emu_pool4(ch x, int n, int H, int W, int C, const float* w, const float* b, mh out) {
  L(dim3(emu_blocks((size_t)n * (H / 4) * (W / 4) * C)), 256, false, pp_pool4_kernel(x, n, H, W, C, w, b, out));
}
void emu_window_attention(ch q, ch k, ch v, ch kp, ch vp, int T, int Hn, int Wn, int C, int ph, int pw, const int* valid, int nvalid, const int* tind, int ntind,
                          const int* masked, mh out) {
  L(dim3((Hn / PP_WH) * (Wn / PP_WW), C / 128, T), 192, true,
    pp_window_attention_kernel(q, k, v, kp, vp, T, Hn, Wn, C, ph, pw, valid, nvalid, tind, ntind, masked, out));
}
void emu_pred_to_rgb8(ch x, int cp, size_t px, uint8_t* out) { L(dim3(emu_blocks(px)), 256, false, pp_pred_to_rgb8_kernel(x, cp, px, out)); }
}
