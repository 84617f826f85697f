// engine.cu - the UNet denoiser engine behind dpir_create / dpir_load_weight / dpir_unet_forward / dpir_unet_x0.
//
// Replays guided_diffusion/unet.py:479-618 (UNetModel.__init__) from dpir_unet_config to get the block structure, packs
// the checkpoint tensors into kernel-native layouts inside one device arena (so a single NCCL broadcast moves the
// model), and turns UNetModel.forward (unet.py:634-663) into a flat "tape" of kernel launches per
// batch shape: activations live in one workspace carved by a lifetime-aware pool at plan time; no allocation,
// synchronisation or host-device copy happens inside a forward.
#include "unet_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace dpir;

namespace {

constexpr size_t ALIGN = 256;
static inline size_t align_up(size_t v, size_t a = ALIGN) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------- structure

struct LayerDesc {
  int kind;  // 0 conv(stem), 1 res, 2 attn
  std::string prefix;
  int cin = 0, cout = 0, updown = 0 /*RS_*/, heads = 0;
};

// how a checkpoint tensor is repacked
enum PackKind { PK_COPY, PK_CONV_STEM, PK_CONV, PK_FILM_W, PK_FILM_B };
struct ParamSlot {
  PackKind kind;
  std::vector<int64_t> shape;
  size_t off = 0;      // byte offset in the arena
  size_t bytes = 0;    // packed bytes
  int ldw = 0;         // padded Cout (fp32: %4, fp16: 16 or %128)
  int film_row = 0;    // PK_FILM_*: first row in the concatenated FiLM matrix
  size_t scale_off = 0; // PK_CONV in split-precision mode: arena offset of the float accumulator scale (1 / weight scale)
  bool loaded = false;
};

struct ConvW { size_t w_off = 0, b_off = 0, s_off = 0; int cin = 0, cout = 0, ldw = 0, ksize = 3; };   // s_off: accumulator scale (split mode)
struct NormW { size_t g_off = 0, b_off = 0; };
struct ResW { NormW n1, n2; ConvW c1, c2, skip; bool has_skip = false; int film_row = 0; };
struct AttnW { NormW n; ConvW qkv, proj; };

struct Buf {   // workspace tensor
  size_t off = 0, bytes = 0;
  int B = 0, H = 0, W = 0, C = 0;
  bool half = false;
  int refs = 0;
  int id = -1;
  size_t stats_off = 0;   // quad statistics slot (workspace offset), valid when has_stats
  bool has_stats = false;
};

enum OpKind { OP_MEMSET, OP_TIME_EMBED, OP_FILM, OP_CONV_IN, OP_GN_STATS, OP_GN_APPLY, OP_CONV, OP_ATTN, OP_GN_COEF, OP_SPLIT_CAST,
              OP_PACK_STEM };
struct Op {
  OpKind kind;
  // generic pointers / ints, interpreted per kind
  GnApplyArgs gn{};
  GnCoefArgs gc{};                                           // OP_GN_COEF
  GnFuse gf{};                                               // OP_CONV with the GroupNorm+SiLU fused on its operand (v4 kernel)
  const void *p2b = nullptr; int Cin2b = 0;                  // second source of a concatenated fused-skip operand
  ConvEpi epi{};
  const void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr;   // workspace tensors (offset-encoded until the plan is rebased)
  void *q0 = nullptr;                                        // workspace tensor (output)
  const void *w = nullptr;                                   // absolute pointer into the weight arena
  const void *w2 = nullptr;                                  // fused 1x1 skip weights (arena) for the persistent 3x3 kernel
  int Cin2 = 0;                                              // channels of the fused skip operand (p2)
  int Cin_alg = 0;                                           // algorithmic input channels when the operand is zero-padded (stem: 3 of 64)
  int act_c = 0;                                             // operand tensor narrower than the K chunk (stem: 8 channels, TMA zero-fills)
  int film_row = -1;                                         // GN apply with FiLM: first row in the FiLM matrix
  size_t bytes = 0;
  int B = 0, H = 0, W = 0, C1 = 0, C2 = 0, Cin = 0, Cout = 0, ldw = 0, ksize = 0, heads = 0, HW = 0, out_half = 0;
  ConvTcPlan *tc = nullptr;
  AttnTcPlan *atc = nullptr;                                 // OP_ATTN on the tcgen05 kernel (fp16 qkv)
  bool bind_x = false, bind_t = false, bind_out = false;   // patched with call arguments at launch
};

struct Plan {
  int B = 0, H = 0, W = 0;
  std::vector<Op> ops;
  std::vector<ConvTcPlan *> tc_plans;
  std::vector<AttnTcPlan *> attn_plans;
  size_t ws_bytes = 0;
  int final_conv = -1;
  cudaGraphExec_t graph_nchw = nullptr, graph_x0 = nullptr;
  ~Plan() {
    for (auto *p : tc_plans) conv_tc_plan_destroy(p);
    for (auto *p : attn_plans) attn_tc_plan_destroy(p);
    if (graph_nchw) cudaGraphExecDestroy(graph_nchw);
    if (graph_x0) cudaGraphExecDestroy(graph_x0);
  }
};

// first-fit offset pool with coalescing; used only at plan time
struct Pool {
  struct Blk { size_t off, size; bool free; };
  std::vector<Blk> blks;
  size_t peak = 0;
  // Every block is followed by a small odd-sized skew (a multiple of 10 KB that cycles with the allocation count).  The
  // activation tensors of this network are exact powers of two (B=32: 2^28 / 2^29 bytes), and back-to-back placement put a
  // convolution's input, output and residual exactly 2^28 bytes apart: measured 3.3x slower on such a layer (0.46 -> 1.5 ms,
  // profiles/r02_workspace_skew.md) - the streams collide in the same DRAM banks.  The skew breaks the alignment.
  int n_alloc = 0;
  size_t alloc(size_t n) {
    n = align_up(n, 1024) + (size_t)((n_alloc++ % 13) + 1) * 10240;
    for (size_t i = 0; i < blks.size(); ++i) {
      if (blks[i].free && blks[i].size >= n) {
        if (blks[i].size > n) {
          Blk rest{blks[i].off + n, blks[i].size - n, true};
          blks[i].size = n;
          blks.insert(blks.begin() + i + 1, rest);
        }
        blks[i].free = false;
        return blks[i].off;
      }
    }
    size_t off = blks.empty() ? 0 : blks.back().off + blks.back().size;
    if (!blks.empty() && blks.back().free) {   // grow the trailing free block
      off = blks.back().off;
      blks.back().size = n;
      blks.back().free = false;
    } else {
      blks.push_back(Blk{off, n, false});
    }
    peak = std::max(peak, off + n);
    return off;
  }
  void release(size_t off) {
    for (size_t i = 0; i < blks.size(); ++i) {
      if (blks[i].off == off && !blks[i].free) {
        blks[i].free = true;
        if (i + 1 < blks.size() && blks[i + 1].free) { blks[i].size += blks[i + 1].size; blks.erase(blks.begin() + i + 1); }
        if (i > 0 && blks[i - 1].free) { blks[i - 1].size += blks[i].size; blks.erase(blks.begin() + i); }
        return;
      }
    }
  }
};

}  // namespace

struct dpir_engine {
  dpir_unet_config cfg{};
  bool half = false;          // tensor-core path (fp16 operands)
  bool stream_half = false;   // fp16s: the residual stream (block outputs, skip tensors) is stored in fp16 too
  int gn_fuse = 1;            // dpir_set_gn_fusion (per engine): ResBlock GroupNorm+SiLU fused into the conv operand where eligible
  int gn_math = 3;            // producer flavour of the fused kernels: 3 activate-once (default), 1 per-tap fp32, 2 per-tap packed half
  bool split = false;         // fp16x3: split-precision tensor-core mode - conv operands are fp16 (hi, lo) pairs, three MMA passes
                              // x_hi*w_hi + x_lo*w_hi + x_hi*w_lo per convolution, fp32 everywhere else (<= 1e-4 parity mode)
  int ted = 0, ch0 = 0;
  std::vector<std::vector<LayerDesc>> inputs, middle, outputs;
  std::map<std::string, ParamSlot> slots;
  std::map<std::string, ResW> resw;
  std::map<std::string, AttnW> attnw;
  ConvW stem, outconv;
  ConvW stem_tc;              // tensor-core pipelines: the stem as a 3x3 convolution over a 64-channel zero-padded operand
  NormW outnorm;
  size_t te_w0 = 0, te_b0 = 0, te_w2 = 0, te_b2 = 0, film_w = 0, film_b = 0, freqs_off = 0, coef_off = 0;
  int film_rows = 0;
  uint8_t *arena = nullptr;
  size_t arena_bytes = 0;
  uint8_t *ws = nullptr;
  size_t ws_bytes = 0;
  std::map<std::tuple<int, int, int>, std::unique_ptr<Plan>> plans;
  int graph_mode = 0;
  // FiLM table: (scale, shift) of every ResBlock for every integer timestep, computed once from the loaded weights
  // (emb_layers(SiLU(time_embed(timestep_embedding(t)))) depends on t only) -> zero embedding work per forward
  int use_film_table = 1;
  float *film_table = nullptr;     // [diffusion_steps][film_rows]
  bool film_table_valid = false;
  // per-op CUDA-event profiler (bench / roofline accounting): events recorded on the launch stream around every op
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;      // pairs
  std::vector<int> prof_op;                  // tape index per pair
  Plan *prof_plan = nullptr;
  // staging buffers for graph mode (fixed addresses baked into the graph)
  float *g_x = nullptr, *g_out = nullptr;
  int64_t *g_t = nullptr;
  size_t g_cap = 0;
  // clamped copy of the caller's timesteps (table lookups stay inside [0, diffusion_steps) whatever the caller passes)
  int64_t *t_buf = nullptr;
  int t_cap = 0;
  // graphs are captured and replayed on an engine-owned stream (the caller may be on the legacy default stream, which
  // cannot be captured); ev_in / ev_out order it after / before the caller's stream
  cudaStream_t gstream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;

  template <typename T> T *ap(size_t off) const { return reinterpret_cast<T *>(arena + off); }
};

namespace {

// ---------------------------------------------------------------------------------------------- construction
int build_structure(dpir_engine *e) {
  const dpir_unet_config &c = e->cfg;
  DPIR_REQUIRE(c.model_channels > 0 && c.model_channels % 128 == 0, "model_channels must be a positive multiple of 128");
  DPIR_REQUIRE(c.n_channel_mult >= 1 && c.n_channel_mult <= 8, "bad channel_mult");
  DPIR_REQUIRE(c.in_channels >= 1 && c.in_channels <= 4, "in_channels must be 1..4");
  DPIR_REQUIRE(c.out_channels >= 3 && c.out_channels <= 16, "out_channels must be 3..16");
  auto has_att = [&](int ds) {
    for (int i = 0; i < c.n_attention_ds; ++i) if (c.attention_ds[i] == ds) return true;
    return false;
  };
  int nh_up = c.num_heads_upsample == -1 ? c.num_heads : c.num_heads_upsample;
  auto heads = [&](int ch, int nh) { return c.num_head_channels == -1 ? nh : ch / c.num_head_channels; };
  int mc = c.model_channels;
  const int den = c.channel_mult_den > 1 ? c.channel_mult_den : 1;
  auto level_ch = [&](int level) { return (int)((int64_t)c.channel_mult[level] * mc / den); };   // int(mult * model_channels), unet.py:483, 497
  int ch = level_ch(0);
  e->ch0 = ch;
  e->ted = mc * 4;
  std::vector<int> chans;
  auto mk = [](int kind, std::string p, int cin, int cout, int ud = 0, int hd = 0) {
    LayerDesc l; l.kind = kind; l.prefix = std::move(p); l.cin = cin; l.cout = cout; l.updown = ud; l.heads = hd; return l;
  };
  e->inputs.push_back({mk(0, "input_blocks.0.0.", c.in_channels, ch)});
  chans.push_back(ch);
  int ds = 1;
  for (int level = 0; level < c.n_channel_mult; ++level) {
    const int lch = level_ch(level);
    for (int r = 0; r < c.num_res_blocks; ++r) {
      int i = (int)e->inputs.size();
      std::vector<LayerDesc> blk;
      blk.push_back(mk(1, "input_blocks." + std::to_string(i) + ".0.", ch, lch));
      ch = lch;
      if (has_att(ds)) blk.push_back(mk(2, "input_blocks." + std::to_string(i) + ".1.", ch, ch, 0, heads(ch, c.num_heads)));
      e->inputs.push_back(blk);
      chans.push_back(ch);
    }
    if (level != c.n_channel_mult - 1) {
      int i = (int)e->inputs.size();
      e->inputs.push_back({mk(1, "input_blocks." + std::to_string(i) + ".0.", ch, ch, RS_DOWN)});
      chans.push_back(ch);
      ds *= 2;
    }
  }
  e->middle.push_back({mk(1, "middle_block.0.", ch, ch), mk(2, "middle_block.1.", ch, ch, 0, heads(ch, c.num_heads)),
                       mk(1, "middle_block.2.", ch, ch)});
  for (int level = c.n_channel_mult - 1; level >= 0; --level) {
    const int lch = level_ch(level);
    for (int i = 0; i <= c.num_res_blocks; ++i) {
      int ich = chans.back();
      chans.pop_back();
      int o = (int)e->outputs.size();
      std::string pre = "output_blocks." + std::to_string(o) + ".";
      std::vector<LayerDesc> blk;
      blk.push_back(mk(1, pre + "0.", ch + ich, lch));
      ch = lch;
      if (has_att(ds)) blk.push_back(mk(2, pre + std::to_string(blk.size()) + ".", ch, ch, 0, heads(ch, nh_up)));
      if (level && i == c.num_res_blocks) {
        blk.push_back(mk(1, pre + std::to_string(blk.size()) + ".", ch, ch, RS_UP));
        ds /= 2;
      }
      e->outputs.push_back(blk);
    }
  }
  return DPIR_OK;
}

struct ArenaBuilder {
  size_t cur = 0;
  size_t take(size_t bytes) { size_t o = cur; cur = align_up(cur + bytes); return o; }
};

int conv_ldw(bool half, int cout) {
  if (half) return cout < 16 ? 16 : cout;
  return (cout + 3) / 4 * 4;
}

int build_params(dpir_engine *e) {
  ArenaBuilder ab;
  const bool half = e->half;
  auto add = [&](const std::string &name, PackKind kind, std::vector<int64_t> shape, size_t bytes, int ldw = 0) -> ParamSlot & {
    ParamSlot s;
    s.kind = kind; s.shape = std::move(shape); s.bytes = bytes; s.ldw = ldw;
    if (kind != PK_FILM_W && kind != PK_FILM_B) s.off = ab.take(bytes);
    e->slots[name] = s;
    return e->slots[name];
  };
  auto add_vec = [&](const std::string &name, int n) { return add(name, PK_COPY, {n}, (size_t)n * 4).off; };
  auto add_conv = [&](const std::string &pre, int cin, int cout, int ksize, bool conv1d) {
    ConvW w;
    w.cin = cin; w.cout = cout; w.ksize = ksize; w.ldw = conv_ldw(half, cout);
    size_t bytes = (size_t)ksize * ksize * cin * w.ldw * (half ? 2 : 4) * (e->split ? 3 : 1);
    std::vector<int64_t> shape = conv1d ? std::vector<int64_t>{cout, cin, 1} : std::vector<int64_t>{cout, cin, ksize, ksize};
    ParamSlot &ws = add(pre + "weight", PK_CONV, shape, bytes, w.ldw);
    w.w_off = ws.off;
    if (e->split) { w.s_off = ab.take(sizeof(float)); e->slots[pre + "weight"].scale_off = w.s_off; }
    w.b_off = add(pre + "bias", PK_COPY, {cout}, (size_t)conv_ldw(false, cout) * 4).off;
    return w;
  };
  int ted = e->ted, mc = e->cfg.model_channels;
  e->te_w0 = add("time_embed.0.weight", PK_COPY, {ted, mc}, (size_t)ted * mc * 4).off;
  e->te_b0 = add_vec("time_embed.0.bias", ted);
  e->te_w2 = add("time_embed.2.weight", PK_COPY, {ted, ted}, (size_t)ted * ted * 4).off;
  e->te_b2 = add_vec("time_embed.2.bias", ted);
  // first pass: count FiLM rows
  int film_rows = 0;
  auto each_layer = [&](auto fn) {
    for (auto *grp : {&e->inputs, &e->middle, &e->outputs})
      for (auto &blk : *grp)
        for (auto &l : blk) fn(l);
  };
  each_layer([&](LayerDesc &l) { if (l.kind == 1) film_rows += 2 * l.cout; });
  e->film_rows = film_rows;
  e->film_w = ab.take((size_t)film_rows * ted * 4);
  e->film_b = ab.take((size_t)film_rows * 4);
  int row = 0;
  int rc = DPIR_OK;
  each_layer([&](LayerDesc &l) {
    const std::string &p = l.prefix;
    if (l.kind == 0) {
      ConvW w;
      w.cin = l.cin; w.cout = l.cout; w.ksize = 3; w.ldw = l.cout;
      w.w_off = add(p + "weight", PK_CONV_STEM, {l.cout, l.cin, 3, 3}, (size_t)9 * l.cin * l.cout * 4, l.cout).off;
      w.b_off = add_vec(p + "bias", l.cout);
      e->stem = w;
      if (half && !e->split && l.cout % 128 == 0) {          // fp16 [tap][cout][64] copy, zero beyond the real input channels
        e->stem_tc = w;
        e->stem_tc.cin = 64; e->stem_tc.ldw = l.cout;
        e->stem_tc.w_off = ab.take((size_t)9 * l.cout * 64 * 2);
      }
    } else if (l.kind == 1) {
      if (l.cin % 128 || l.cout % 128) { set_error("channel counts must be multiples of 128 (%s)", p.c_str()); rc = DPIR_E_ARG; return; }
      ResW r;
      r.n1.g_off = add_vec(p + "in_layers.0.weight", l.cin);
      r.n1.b_off = add_vec(p + "in_layers.0.bias", l.cin);
      r.c1 = add_conv(p + "in_layers.2.", l.cin, l.cout, 3, false);
      r.film_row = row;
      ParamSlot &fw = add(p + "emb_layers.1.weight", PK_FILM_W, {2 * l.cout, ted}, (size_t)2 * l.cout * ted * 4);
      fw.film_row = row; fw.off = e->film_w + (size_t)row * ted * 4;
      ParamSlot &fb = add(p + "emb_layers.1.bias", PK_FILM_B, {2 * l.cout}, (size_t)2 * l.cout * 4);
      fb.film_row = row; fb.off = e->film_b + (size_t)row * 4;
      row += 2 * l.cout;
      r.n2.g_off = add_vec(p + "out_layers.0.weight", l.cout);
      r.n2.b_off = add_vec(p + "out_layers.0.bias", l.cout);
      r.c2 = add_conv(p + "out_layers.3.", l.cout, l.cout, 3, false);
      r.has_skip = l.cin != l.cout;
      if (r.has_skip) r.skip = add_conv(p + "skip_connection.", l.cin, l.cout, 1, false);
      e->resw[p] = r;
    } else {
      if (l.cin % 128 || l.cin != l.heads * 64) { set_error("attention needs C %% 128 == 0 and head dim 64 (%s)", p.c_str()); rc = DPIR_E_ARG; return; }
      AttnW a;
      a.n.g_off = add_vec(p + "norm.weight", l.cin);
      a.n.b_off = add_vec(p + "norm.bias", l.cin);
      a.qkv = add_conv(p + "qkv.", l.cin, 3 * l.cin, 1, true);
      a.proj = add_conv(p + "proj_out.", l.cin, l.cin, 1, true);
      e->attnw[p] = a;
    }
  });
  if (rc) return rc;
  e->outnorm.g_off = add_vec("out.0.weight", e->ch0);
  e->outnorm.b_off = add_vec("out.0.bias", e->ch0);
  e->outconv = add_conv("out.2.", e->ch0, e->cfg.out_channels, 3, false);
  e->freqs_off = ab.take((size_t)(mc / 2) * 4);
  e->coef_off = ab.take((size_t)e->cfg.diffusion_steps * 2 * 4);
  e->arena_bytes = ab.cur;
  return DPIR_OK;
}

int init_tables(dpir_engine *e) {
  // nn.py:112-114 frequencies and gaussian_diffusion.py:133-151 eps->x0 coefficients (float64 tables -> float)
  int half = e->cfg.model_channels / 2;
  std::vector<float> fr(half);
  for (int i = 0; i < half; ++i) fr[i] = expf(-logf(10000.0f) * (float)i / (float)half);
  DPIR_CUDA_CHECK(cudaMemcpy(e->arena + e->freqs_off, fr.data(), fr.size() * 4, cudaMemcpyHostToDevice));
  int T = e->cfg.diffusion_steps;
  std::vector<float> coef((size_t)T * 2);
  double scale = 1000.0 / T, b0 = scale * 0.0001, b1 = scale * 0.02, ac = 1.0;
  for (int i = 0; i < T; ++i) {
    double beta = T > 1 ? b0 + (b1 - b0) * (double)i / (double)(T - 1) : b0;
    ac *= (1.0 - beta);
    coef[2 * i] = (float)std::sqrt(1.0 / ac);
    coef[2 * i + 1] = (float)std::sqrt(1.0 / ac - 1.0);
  }
  DPIR_CUDA_CHECK(cudaMemcpy(e->arena + e->coef_off, coef.data(), coef.size() * 4, cudaMemcpyHostToDevice));
  return DPIR_OK;
}

// ---------------------------------------------------------------------------------------------- plan building
struct PlanBuilder {
  dpir_engine *e;
  Plan *plan;
  Pool pool;
  std::vector<Buf> bufs;
  size_t stats_off = 0;      // inside the workspace: [n_norms][B][64] doubles
  int n_norms = 0;
  size_t film_off = 0, hidden_off = 0, emb_off = 0;
  int B, H, W;
  int err = DPIR_OK;

  size_t stats_cur = 0, stats_bytes = 0;

  int new_buf(int B_, int H_, int W_, int C_, bool half, bool want_stats = false) {
    Buf b;
    b.B = B_; b.H = H_; b.W = W_; b.C = C_; b.half = half;
    b.bytes = (size_t)B_ * H_ * W_ * C_ * (half ? 2 : 4);
    b.off = pool.alloc(b.bytes);
    b.refs = 1;
    b.id = (int)bufs.size();
    if (want_stats) {
      size_t need = (size_t)B_ * (C_ / 4) * 2 * sizeof(double);
      if (stats_cur + need > stats_bytes) { set_error("internal: statistics arena overflow"); err = DPIR_E_STATE; }
      b.stats_off = stats_off + stats_cur;
      b.has_stats = true;
      stats_cur += need;
    }
    bufs.push_back(b);
    return b.id;
  }
  void retain(int id) { bufs[id].refs++; }
  void release(int id) {
    if (id < 0) return;
    if (--bufs[id].refs == 0) pool.release(bufs[id].off);
  }
  // pointers are stored as workspace offsets (+1 so that 0 stays "null") and rebased after the workspace is allocated
  template <typename T> T *wp(size_t off) { return reinterpret_cast<T *>(off + 1); }
  template <typename T> T *bp(int id) { return wp<T>(bufs[id].off); }
  double *sp(int id) {
    if (!bufs[id].has_stats) { set_error("internal: tensor %d has no statistics", id); err = DPIR_E_STATE; return nullptr; }
    return wp<double>(bufs[id].stats_off);
  }

  // separate statistics pass over an fp32 tensor (stem output; every conv output in fp32 mode)
  void emit_stats(int id) {
    const Buf &t = bufs[id];
    Op st;
    st.kind = OP_GN_STATS;
    st.p0 = bp<float>(id); st.C1 = t.C; st.out_half = t.half ? 1 : 0;
    st.q0 = sp(id); st.B = B; st.HW = t.H * t.W;
    plan->ops.push_back(st);
  }

  int gn(int src1, int src2, const NormW &nw, int film_row /*-1: none*/, int silu, int resample, bool want_raw32,
         bool want_rawact, int *act_id, int *raw32_id, int *rawact_id) {
    const Buf s1 = bufs[src1];   // by value: new_buf() below may reallocate `bufs`
    int C1 = s1.C, C2 = src2 >= 0 ? bufs[src2].C : 0, C = C1 + C2;
    int Hin = s1.H, Win = s1.W;
    int Ho = resample == RS_DOWN ? Hin / 2 : (resample == RS_UP ? Hin * 2 : Hin);
    int Wo = resample == RS_DOWN ? Win / 2 : (resample == RS_UP ? Win * 2 : Win);
    n_norms++;
    // fp16 stream: the raw (resampled / concatenated) residual copy is fp16 as well and is written through `rawact`
    const bool raw_as_act = want_raw32 && e->stream_half;
    const int Cop = e->split ? 2 * C : C;            // operand tensors hold [hi | lo] pairs in split-precision mode
    *act_id = new_buf(B, Ho, Wo, Cop, e->half);
    *raw32_id = (want_raw32 || raw_as_act) ? new_buf(B, Ho, Wo, C, raw_as_act) : -1;
    *rawact_id = want_rawact ? new_buf(B, Ho, Wo, Cop, e->half) : -1;
    if (raw_as_act && want_rawact) { set_error("internal: raw copy requested twice"); err = DPIR_E_STATE; }
    Op ap;
    ap.kind = OP_GN_APPLY;
    GnApplyArgs &a = ap.gn;
    a.src1 = bp<void>(src1); a.C1 = C1;
    a.src2 = src2 >= 0 ? bp<void>(src2) : nullptr; a.C2 = C2;
    a.src_half = bufs[src1].half ? 1 : 0;
    a.q1 = sp(src1);
    a.q2 = src2 >= 0 ? sp(src2) : nullptr;
    a.gamma = e->ap<float>(nw.g_off); a.beta = e->ap<float>(nw.b_off);
    a.film = film_row >= 0 ? wp<float>(film_off + (size_t)film_row * 4) : nullptr;
    a.film_stride = e->film_rows;
    ap.film_row = film_row;
    a.silu = silu; a.resample = resample; a.B = B; a.H = Hin; a.W = Win; a.out_half = e->half; a.split = e->split ? 1 : 0;
    a.act = bp<void>(*act_id);
    a.raw32 = (want_raw32 && !raw_as_act) ? bp<float>(*raw32_id) : nullptr;
    a.rawact = want_rawact ? bp<void>(*rawact_id) : (raw_as_act ? bp<void>(*raw32_id) : nullptr);
    plan->ops.push_back(ap);
    return DPIR_OK;
  }

  // conv on a workspace activation -> new NHWC buffer (fp32, or fp16 when out_half in tensor-core mode); residual
  // optional; want_stats attaches quad statistics to the output (fused in the tcgen05 epilogue, separate pass in fp32 mode)
  // whether the tensor-core path will run this conv on the persistent 3x3 kernel (mirrors conv_tc_plan_create)
  bool v2_eligible(int act, const ConvW &cw) const {
    const Buf &a = bufs[act];
    return e->half && conv_tc_v2_enabled() && cw.ksize == 3 && a.H >= 16 && a.W >= 16 && cw.ldw % 128 == 0;
  }

  int conv(int act, const ConvW &cw, int residual, int *out_id, bool want_stats = false, bool out_half = false,
           int skip_act = -1, const ConvW *skip_w = nullptr, int skip_act2 = -1, bool up = false) {
    Buf a = bufs[act];   // by value: new_buf() below may reallocate `bufs`
    if (up) { a.H *= 2; a.W *= 2; }                    // `act` is the half-resolution source of an up-sampled (index-mapped) operand
    out_half = out_half && e->half && !e->split;       // split-precision mode keeps every conv output in fp32
    *out_id = new_buf(B, a.H, a.W, cw.cout, out_half, want_stats);
    if (residual >= 0 && bufs[residual].half && !e->half) { set_error("internal: fp16 residual in fp32 mode"); err = DPIR_E_STATE; }
    Op op;
    op.kind = OP_CONV;
    op.p0 = bp<void>(act);
    op.w = e->arena + cw.w_off;
    op.B = B; op.H = a.H; op.W = a.W; op.Cin = cw.cin; op.Cout = cw.cout; op.ldw = cw.ldw; op.ksize = cw.ksize;
    op.epi.mode = EPI_NHWC;
    op.epi.bias = e->ap<float>(cw.b_off);
    op.epi.residual = residual >= 0 ? bp<float>(residual) : nullptr;
    op.epi.res_half = (residual >= 0 && bufs[residual].half) ? 1 : 0;
    op.epi.out = bp<float>(*out_id);
    op.epi.out_half = out_half ? 1 : 0;
    op.epi.qstats = (want_stats && e->half) ? sp(*out_id) : nullptr;
    op.epi.acc_scale = e->split ? e->ap<float>(cw.s_off) : nullptr;
    if (skip_act >= 0) {   // out = conv3x3(act) + conv1x1(skip_act): one accumulation (unet.py:222, 256)
      op.p2 = bp<void>(skip_act);
      op.w2 = e->arena + skip_w->w_off;
      op.Cin2 = skip_w->cin;
      if (skip_act2 >= 0) {          // the skip operand is the virtual concat (skip_act | skip_act2)
        op.p2b = bp<void>(skip_act2);
        op.Cin2b = bufs[skip_act2].C;
        if (bufs[skip_act].C + op.Cin2b != skip_w->cin) { set_error("internal: concat skip channels"); err = DPIR_E_STATE; }
      }
      op.epi.bias2 = e->ap<float>(skip_w->b_off);
    }
    plan->ops.push_back(op);
    if (want_stats && !e->half) emit_stats(*out_id);
    return DPIR_OK;
  }

  // v4: instead of a k_gn_apply pass, emit the tiny per-(image, channel) coefficient kernel; the conv applies
  // GroupNorm(+FiLM)+SiLU to its operand on the fly.  Returns the coefficient buffer id (release it after the conv).
  int gn_coef(int src1, int src2, const NormW &nw, int film_row) {
    const Buf s1 = bufs[src1];
    const int C1 = s1.C, C2 = src2 >= 0 ? bufs[src2].C : 0, C = C1 + C2;
    n_norms++;
    int cb = new_buf(B, 1, 1, 2 * C, false);
    Op op;
    op.kind = OP_GN_COEF;
    GnCoefArgs &a = op.gc;
    a.C1 = C1; a.C2 = C2;
    a.q1 = sp(src1); a.q2 = src2 >= 0 ? sp(src2) : nullptr;
    a.gamma = e->ap<float>(nw.g_off); a.beta = e->ap<float>(nw.b_off);
    a.film = film_row >= 0 ? wp<float>(film_off + (size_t)film_row * 4) : nullptr;
    a.film_stride = e->film_rows; a.tsteps = nullptr;
    op.film_row = film_row;
    a.B = B; a.HW = s1.H * s1.W;
    a.coef = bp<float>(cb);
    op.H = s1.H; op.W = s1.W; op.Cin = C;
    plan->ops.push_back(op);
    return cb;
  }
  // the conv just pushed reads SiLU(GroupNorm(src1|src2)) instead of a materialised activation
  void fuse_last_conv(int src1, int src2, int coef_buf) {
    Op &op = plan->ops.back();
    op.gf.src1 = bp<__half>(src1); op.gf.C1 = bufs[src1].C;
    op.gf.src2 = src2 >= 0 ? bp<__half>(src2) : nullptr; op.gf.C2 = src2 >= 0 ? bufs[src2].C : 0;
    op.gf.coef = bp<float>(coef_buf);
    op.p0 = op.gf.src1;
  }
  // ResBlock eligible for the fused-operand kernel: fp16 stream, no resampling, 3x3 convs on the persistent swap-AB kernel
  bool gn_fusable(const LayerDesc &l, int x, int skip, const ResW &rw) const {
    const Buf &a = bufs[x];
    if (!(e->gn_fuse && e->stream_half && conv_tc_swapab_enabled() && l.updown == RS_NONE && a.H >= 16 && a.W >= 16)) return false;
    if (!a.half || (skip >= 0 && !bufs[skip].half)) return false;
    if (a.C % 64 || (skip >= 0 && bufs[skip].C % 64)) return false;
    if (rw.c1.ksize != 3 || rw.c2.ksize != 3 || rw.c1.ldw % 128 || rw.c2.ldw % 128 || rw.c1.cout % 64) return false;
    if (skip >= 0 && !rw.has_skip) return false;     // a concat input without a 1x1 skip would need the raw concat copy
    return true;
  }

  // Up-sampling ResBlock (unet.py:237-244, 256: h = conv1(up(SiLU(GN(x)))), out = up(x) + conv2(...)) without the up-sampled
  // tensors: conv1's producers read x through the nearest-neighbour index map (GroupNorm statistics of up(x) = those of x), conv2's
  // epilogue reads its residual up(x) from x the same way.  Saves the k_gn_apply pass that wrote up(SiLU(GN(x))) and up(x) at 4x
  // the pixels.  fp16 stream + activate-once producers only.
  bool up_fusable(const LayerDesc &l, int x, int skip, const ResW &rw) const {
    const Buf &a = bufs[x];
    static int up_fuse = -1;           // DPIR_UP_FUSE=0 switches it off (A/B timing)
    if (up_fuse < 0) { const char *ev = getenv("DPIR_UP_FUSE"); up_fuse = (ev && ev[0] == '0') ? 0 : 1; }
    return up_fuse && e->gn_fuse && e->gn_math == 3 && e->stream_half && conv_tc_swapab_enabled() && l.updown == RS_UP && skip < 0 &&
           !rw.has_skip && a.half && a.H >= 8 && a.W >= 8 && a.C % 64 == 0 && rw.c1.ksize == 3 && rw.c2.ksize == 3 &&
           rw.c1.ldw % 128 == 0 && rw.c2.ldw % 128 == 0 && rw.c1.cout % 64 == 0;
  }
  int resblock_up_fused(const LayerDesc &l, int x, int *out) {
    const ResW &rw = e->resw[l.prefix];
    int c1 = gn_coef(x, -1, rw.n1, -1);
    int h1;
    conv(x, rw.c1, -1, &h1, /*want_stats=*/true, /*out_half=*/true, -1, nullptr, -1, /*up=*/true);
    fuse_last_conv(x, -1, c1);
    plan->ops.back().gf.up = 1;
    release(c1);
    int a2, d0, d1;
    gn(h1, -1, rw.n2, rw.film_row, 1, RS_NONE, false, false, &a2, &d0, &d1);
    release(h1);
    conv(a2, rw.c2, x, out, /*want_stats=*/true, /*out_half=*/true);        // residual = x at half resolution, index-mapped in the epilogue
    plan->ops.back().epi.res_up = 1;
    release(a2);
    release(x);
    return DPIR_OK;
  }

  int resblock_fused(const LayerDesc &l, int x, int skip, int *out) {
    const ResW &rw = e->resw[l.prefix];
    // in_layers: conv1(SiLU(GN(x | skip)))
    int c1 = gn_coef(x, skip, rw.n1, -1);
    int h1;
    conv(x, rw.c1, -1, &h1, /*want_stats=*/true, /*out_half=*/true);
    fuse_last_conv(x, skip, c1);
    release(c1);
    // out_layers: conv2(SiLU(FiLM(GN(h1)))) + skip_connection(x | skip)
    int c2 = gn_coef(h1, -1, rw.n2, rw.film_row);
    if (rw.has_skip) {
      conv(h1, rw.c2, -1, out, /*want_stats=*/true, /*out_half=*/true, x, &rw.skip, skip);
    } else {
      conv(h1, rw.c2, x, out, /*want_stats=*/true, /*out_half=*/true);
    }
    fuse_last_conv(h1, -1, c2);
    release(c2);
    release(h1);
    release(x);
    if (skip >= 0) release(skip);
    return DPIR_OK;
  }

  int resblock(const LayerDesc &l, int x, int skip, int *out) {
    const ResW &rw = e->resw[l.prefix];
    if (gn_fusable(l, x, skip, rw)) return resblock_fused(l, x, skip, out);
    if (up_fusable(l, x, skip, rw)) return resblock_up_fused(l, x, out);
    bool concat = skip >= 0;
    bool need_raw32 = !rw.has_skip && (l.updown != RS_NONE || concat);
    int a1, xr, xraw;
    gn(x, skip, rw.n1, -1, 1, l.updown, need_raw32, rw.has_skip, &a1, &xr, &xraw);
    int residual = -1;
    if (!rw.has_skip) {
      if (need_raw32) { residual = xr; release(x); }
      else residual = x;   // keep x alive until conv2
    } else {
      release(x);
    }
    if (concat) release(skip);
    int h1;
    conv(a1, rw.c1, -1, &h1, /*want_stats=*/true, /*out_half=*/true);   // h1 feeds only GroupNorm -> conv2
    release(a1);
    // (Fusing the SECOND norm of the up/down-sampling ResBlocks into conv2 was measured in round 2: the k_gn_apply pass it removes
    // (0.48 ms per forward) is paid back by the slower fused convolution (+0.59 ms) - profiles/r02_conv_ring_experiments.md - so
    // these blocks keep the separate pass.)
    int a2, d0, d1;
    gn(h1, -1, rw.n2, rw.film_row, 1, RS_NONE, false, false, &a2, &d0, &d1);
    release(h1);
    if (rw.has_skip && v2_eligible(a2, rw.c2) && !e->split) {
      conv(a2, rw.c2, -1, out, /*want_stats=*/true, /*out_half=*/e->stream_half, xraw, &rw.skip);
      release(xraw);
    } else {
      if (rw.has_skip) {
        conv(xraw, rw.skip, -1, &residual);
        release(xraw);
      }
      conv(a2, rw.c2, residual, out, /*want_stats=*/true, /*out_half=*/e->stream_half);
      release(residual);
    }
    release(a2);
    return DPIR_OK;
  }

  int attnblock(const LayerDesc &l, int x, int *out) {
    const AttnW &aw = e->attnw[l.prefix];
    int xn, d0, d1;
    gn(x, -1, aw.n, -1, 0, RS_NONE, false, false, &xn, &d0, &d1);
    int qkv;
    conv(xn, aw.qkv, -1, &qkv, /*want_stats=*/false, /*out_half=*/true);   // fp16 in the tensor-core pipelines: the attention's TMA operand
    release(xn);
    const Buf &xb = bufs[x];
    const int aH = xb.H, aW = xb.W;
    int a = new_buf(B, aH, aW, l.cin, e->half && !e->split);
    Op op;
    op.kind = OP_ATTN;
    op.p0 = bp<float>(qkv); op.q0 = bp<void>(a);
    op.B = B; op.HW = aH * aW; op.Cin = l.cin; op.heads = l.heads; op.out_half = (e->half && !e->split) ? 1 : 0;
    op.C2 = bufs[qkv].half ? 1 : 0;            // qkv is fp16 -> tcgen05 attention
    plan->ops.push_back(op);
    release(qkv);
    if (e->split) {      // fp32 attention output -> [hi | lo] operand of the proj convolution
      int a2 = new_buf(B, aH, aW, 2 * l.cin, true);
      Op sc;
      sc.kind = OP_SPLIT_CAST;
      sc.p0 = bp<float>(a); sc.q0 = bp<void>(a2);
      sc.B = B; sc.HW = aH * aW; sc.Cin = l.cin;
      plan->ops.push_back(sc);
      release(a);
      a = a2;
    }
    conv(a, aw.proj, x, out, /*want_stats=*/true, /*out_half=*/e->stream_half);
    release(a);
    release(x);
    return DPIR_OK;
  }

  int run_block(const std::vector<LayerDesc> &blk, int h, int skip, int *out) {
    for (size_t j = 0; j < blk.size(); ++j) {
      const LayerDesc &l = blk[j];
      int o;
      if (l.kind == 1) resblock(l, h, j == 0 ? skip : -1, &o);
      else attnblock(l, h, &o);
      h = o;
    }
    *out = h;
    return DPIR_OK;
  }

  int build() {
    int ted = e->ted;
    // persistent small buffers first (never released)
    hidden_off = pool.alloc((size_t)B * ted * 4);
    emb_off = pool.alloc((size_t)B * ted * 4);
    film_off = pool.alloc((size_t)B * e->film_rows * 4);
    int total_norms = 1;
    size_t total_c = e->ch0;
    for (auto *grp : {&e->inputs, &e->middle, &e->outputs})
      for (auto &blk : *grp)
        for (auto &l : blk) {
          total_norms += l.kind == 1 ? 2 : (l.kind == 2 ? 1 : 0);
          total_c += l.kind == 1 ? 2 * (size_t)l.cout : (l.kind == 2 ? (size_t)l.cin : 0);
        }
    stats_bytes = (size_t)B * (total_c / 4) * 2 * sizeof(double);
    stats_off = pool.alloc(stats_bytes);

    Op ms; ms.kind = OP_MEMSET; ms.q0 = wp<void>(stats_off); ms.bytes = stats_bytes;
    plan->ops.push_back(ms);
    Op te; te.kind = OP_TIME_EMBED; te.bind_t = true; te.q0 = wp<void>(hidden_off); te.p1 = wp<void>(emb_off); te.B = B;
    plan->ops.push_back(te);
    Op fl; fl.kind = OP_FILM; fl.p0 = wp<void>(emb_off); fl.q0 = wp<void>(film_off); fl.B = B;
    plan->ops.push_back(fl);

    std::vector<int> hs;
    int h;
    static int stem_simt = -1;           // DPIR_STEM_SIMT=1: keep the CUDA-core stem in the tensor-core pipelines (A/B timing)
    if (stem_simt < 0) { const char *ev = getenv("DPIR_STEM_SIMT"); stem_simt = (ev && ev[0] == '1') ? 1 : 0; }
    if (e->stem_tc.cin && !stem_simt && H >= 16 && W >= 16) {
      // tensor-core stem: repack x as an 8-channel fp16 NHWC operand (the conv's TMA zero-fills the rest of the 64-channel K chunk),
      // then the ordinary 3x3 path (bias, store, GroupNorm statistics in the tcgen05 epilogue - no separate statistics pass)
      int xp = new_buf(B, H, W, 8, true);
      Op pk; pk.kind = OP_PACK_STEM; pk.bind_x = true; pk.q0 = bp<void>(xp); pk.B = B; pk.H = H; pk.W = W; pk.Cin = e->stem.cin;
      plan->ops.push_back(pk);
      conv(xp, e->stem_tc, -1, &h, /*want_stats=*/true, /*out_half=*/e->stream_half);
      plan->ops.back().Cin_alg = e->stem.cin;
      plan->ops.back().act_c = 8;
      release(xp);
    } else {
      h = new_buf(B, H, W, e->ch0, e->stream_half, /*want_stats=*/true);
      Op ci; ci.kind = OP_CONV_IN; ci.bind_x = true; ci.q0 = bp<void>(h); ci.B = B; ci.H = H; ci.W = W;
      ci.out_half = e->stream_half ? 1 : 0;
      plan->ops.push_back(ci);
      emit_stats(h);
    }
    retain(h); hs.push_back(h);
    for (size_t i = 1; i < e->inputs.size(); ++i) {
      int o;
      run_block(e->inputs[i], h, -1, &o);
      h = o;
      retain(h); hs.push_back(h);
    }
    { int o; run_block(e->middle[0], h, -1, &o); h = o; }
    for (size_t i = 0; i < e->outputs.size(); ++i) {
      int skip = hs.back(); hs.pop_back();
      int o;
      run_block(e->outputs[i], h, skip, &o);
      h = o;
    }
    // out.0-2 (unet.py:617-621): GroupNorm + SiLU + conv3x3 -> 6 channels.  fp16 stream: the norm is applied inside the narrow-output
    // kernel (k_conv3x3_dxn<true>: the TMA fetches the raw h, activator warps rewrite the halo boxes in place); DPIR_OUT_FUSE=0 keeps
    // the separate k_gn_apply pass (A/B timing)
    static int out_fuse = -1;
    if (out_fuse < 0) { const char *ev = getenv("DPIR_OUT_FUSE"); out_fuse = (ev && ev[0] == '0') ? 0 : 1; }
    const bool fuse_out = out_fuse && e->gn_fuse && e->stream_half && !e->split && bufs[h].half && conv_tc_swapab_enabled() &&
                          conv_tc_narrow_eligible(H, W, e->outconv.cin, e->outconv.ldw);
    int a = -1, d0, d1, cb = -1;
    if (fuse_out) cb = gn_coef(h, -1, e->outnorm, -1);
    else { gn(h, -1, e->outnorm, -1, 1, RS_NONE, false, false, &a, &d0, &d1); release(h); }
    Op fc;
    fc.kind = OP_CONV;
    if (fuse_out) { fc.gf.src1 = bp<__half>(h); fc.gf.C1 = bufs[h].C; fc.gf.coef = bp<float>(cb); }
    fc.p0 = fuse_out ? bp<void>(h) : bp<void>(a); fc.w = e->arena + e->outconv.w_off;
    fc.B = B; fc.H = H; fc.W = W; fc.Cin = e->outconv.cin; fc.Cout = e->outconv.cout; fc.ldw = e->outconv.ldw; fc.ksize = 3;
    fc.epi.mode = EPI_NCHW; fc.epi.bias = e->ap<float>(e->outconv.b_off);
    fc.epi.coef = e->ap<float>(e->coef_off);
    fc.epi.coef_rows = e->cfg.diffusion_steps;
    fc.epi.acc_scale = e->split ? e->ap<float>(e->outconv.s_off) : nullptr;
    fc.bind_out = true;
    plan->final_conv = (int)plan->ops.size();
    plan->ops.push_back(fc);
    if (fuse_out) { release(cb); release(h); }
    else release(a);
    if (n_norms != total_norms) { set_error("internal: norm count mismatch %d vs %d", n_norms, total_norms); return DPIR_E_STATE; }
    plan->ws_bytes = pool.peak;
    return err;
  }
};

template <typename T> void rebase(T *&p, uint8_t *base) {
  if (p) p = reinterpret_cast<T *>(base + (reinterpret_cast<uintptr_t>(p) - 1));
}
void rebase_c(const void *&p, uint8_t *base) {
  if (p) p = base + (reinterpret_cast<uintptr_t>(p) - 1);
}

int get_plan(dpir_engine *e, int B, int H, int W, Plan **out) {
  auto key = std::make_tuple(B, H, W);
  auto it = e->plans.find(key);
  if (it != e->plans.end()) { *out = it->second.get(); return DPIR_OK; }
  int nlev = e->cfg.n_channel_mult;
  DPIR_REQUIRE(B >= 1 && H >= 1 && W >= 1, "bad batch shape");
  DPIR_REQUIRE(H % (1 << (nlev - 1)) == 0 && W % (1 << (nlev - 1)) == 0, "H and W must be divisible by %d", 1 << (nlev - 1));
  std::unique_ptr<Plan> plan(new Plan());
  plan->B = B; plan->H = H; plan->W = W;
  PlanBuilder pb;
  pb.e = e; pb.plan = plan.get(); pb.B = B; pb.H = H; pb.W = W;
  int rc = pb.build();
  if (rc) return rc;
  if (plan->ws_bytes > e->ws_bytes) {
    // grow the workspace: cached plans hold absolute addresses (and TMA descriptors), so drop them
    e->plans.clear();
    if (e->ws) { cudaDeviceSynchronize(); cudaFree(e->ws); e->ws = nullptr; e->ws_bytes = 0; }
    cudaError_t ce = cudaMalloc(&e->ws, plan->ws_bytes);
    if (ce != cudaSuccess) { set_error("workspace cudaMalloc(%zu) failed: %s", plan->ws_bytes, cudaGetErrorString(ce)); return DPIR_E_NOMEM; }
    e->ws_bytes = plan->ws_bytes;
  }
  uint8_t *base = e->ws;
  for (Op &op : plan->ops) {
    rebase_c(op.p0, base); rebase_c(op.p1, base); rebase_c(op.p2, base); rebase(op.q0, base);
    if (op.kind == OP_GN_APPLY) {
      rebase_c(op.gn.src1, base); rebase_c(op.gn.src2, base); rebase(op.gn.q1, base); rebase(op.gn.q2, base);
      rebase(op.gn.film, base);
      rebase(op.gn.act, base); rebase(op.gn.raw32, base); rebase(op.gn.rawact, base);
    }
    if (op.kind == OP_CONV) {
      rebase(op.epi.residual, base); rebase(op.epi.out, base); rebase(op.epi.qstats, base);
      rebase_c(op.p2b, base);
      rebase(op.gf.src1, base); rebase(op.gf.src2, base); rebase(op.gf.coef, base);
    }
    if (op.kind == OP_GN_COEF) { rebase(op.gc.q1, base); rebase(op.gc.q2, base); rebase(op.gc.film, base); rebase(op.gc.coef, base); }
  }
  for (Op &op : plan->ops) {
    if (op.kind == OP_ATTN && op.C2) {
      AttnTcPlan *ap = nullptr;
      rc = attn_tc_plan_create(op.p0, op.B, op.HW, op.Cin, &ap);
      if (rc) return rc;
      op.atc = ap;
      plan->attn_plans.push_back(ap);
    }
  }
  if (e->half) {
    for (Op &op : plan->ops) {
      if (op.kind != OP_CONV) continue;
      ConvTcPlan *tp = nullptr;
      rc = conv_tc_plan_create(op.p0, op.w, op.B, op.H, op.W, op.Cin, op.ldw, op.ksize, &tp, e->split ? 1 : 0, 0, op.act_c);
      if (rc) return rc;
      op.tc = tp;
      plan->tc_plans.push_back(tp);
      if (conv_tc_plan_is_v2(tp) && !op.bind_out) {
        rc = conv_tc_plan_set_io(tp, op.epi.out, op.epi.out_half, op.epi.residual, op.epi.res_half, op.Cout, op.epi.res_up);
        if (rc) return rc;
      }
      if (op.p2) {
        if (!conv_tc_plan_is_v2(tp)) { set_error("internal: skip fusion planned for a non-persistent conv"); return DPIR_E_STATE; }
        rc = conv_tc_plan_add_skip(tp, op.p2, op.w2, op.Cin2 - op.Cin2b, op.p2b, op.Cin2b);
        if (rc) return rc;
      }
      if (op.gf.coef) {
        rc = conv_tc_plan_set_gn_fuse(tp, op.gf, e->gn_math);
        if (rc) return rc;
      }
    }
  }
  *out = plan.get();
  e->plans[key] = std::move(plan);
  return DPIR_OK;
}

int run_plan(dpir_engine *e, Plan *plan, const float *x, const int64_t *t, float *out, int mode, cudaStream_t st) {
  int op_index = -1;
  for (Op &op : plan->ops) {
    int rc = DPIR_OK;
    ++op_index;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (e->profiling) {
      DPIR_CUDA_CHECK(cudaEventCreate(&ev0));
      DPIR_CUDA_CHECK(cudaEventCreate(&ev1));
      DPIR_CUDA_CHECK(cudaEventRecord(ev0, st));
    }
    switch (op.kind) {
      case OP_MEMSET:
        DPIR_CUDA_CHECK(cudaMemsetAsync(op.q0, 0, op.bytes, st));
        break;
      case OP_TIME_EMBED:
        if (e->use_film_table) break;
        rc = launch_time_embed(t, e->ap<float>(e->freqs_off), e->ap<float>(e->te_w0), e->ap<float>(e->te_b0),
                               e->ap<float>(e->te_w2), e->ap<float>(e->te_b2), (float *)op.q0, (float *)const_cast<void *>(op.p1),
                               op.B, e->cfg.model_channels, e->ted, st);
        break;
      case OP_FILM:
        if (e->use_film_table) break;
        rc = launch_linear_rows((const float *)op.p0, e->ap<float>(e->film_w), e->ap<float>(e->film_b), (float *)op.q0, op.B,
                                e->ted, e->film_rows, 0, 0, st);
        break;
      case OP_CONV_IN:
        rc = launch_conv_in(x, e->ap<float>(e->stem.w_off), e->ap<float>(e->stem.b_off), op.q0, op.out_half, op.B, op.H, op.W,
                            e->stem.cin, e->stem.cout, st);
        break;
      case OP_PACK_STEM:
        rc = launch_pack_stem(x, (__half *)op.q0, op.B, op.Cin, op.H, op.W, st);
        break;
      case OP_GN_STATS:
        rc = launch_gn_stats(op.p0, op.out_half, op.C1, (double *)op.q0, op.B, op.HW, st);
        break;
      case OP_GN_APPLY:
        if (op.film_row >= 0 && e->use_film_table) {
          GnApplyArgs ga = op.gn;
          ga.film = e->film_table + op.film_row;
          ga.tsteps = t;
          rc = launch_gn_apply(ga, st);
        } else {
          rc = launch_gn_apply(op.gn, st);
        }
        break;
      case OP_GN_COEF: {
        GnCoefArgs gc = op.gc;
        if (op.film_row >= 0 && e->use_film_table) { gc.film = e->film_table + op.film_row; gc.tsteps = t; }
        rc = launch_gn_coef(gc, st);
        break;
      }
      case OP_CONV: {
        ConvEpi epi = op.epi;
        if (op.bind_out) {
          epi.mode = mode;
          epi.out = out;
          epi.x_nchw = x;
          epi.t = t;
        }
        if (e->half) rc = launch_conv_tc(op.tc, epi, op.Cout, st);
        else rc = launch_conv_simt((const float *)op.p0, (const float *)op.w, op.ldw, epi, op.B, op.H, op.W, op.Cin, op.Cout,
                                   op.ksize, st);
        break;
      }
      case OP_SPLIT_CAST:
        rc = launch_split_cast((const float *)op.p0, (__half *)op.q0, (int64_t)op.B * op.HW, op.Cin, st);
        break;
      case OP_ATTN:
        if (op.atc) rc = launch_attention_tc(op.atc, op.q0, op.out_half, op.heads, e->cfg.use_new_attention_order, st);
        else rc = launch_attention((const float *)op.p0, op.q0, op.out_half, op.B, op.HW, op.Cin, op.heads,
                                   e->cfg.use_new_attention_order, st);
        break;
    }
    if (e->profiling) {
      DPIR_CUDA_CHECK(cudaEventRecord(ev1, st));
      e->prof_events.push_back(ev0);
      e->prof_events.push_back(ev1);
      e->prof_op.push_back(op_index);
      e->prof_plan = plan;
    }
    if (rc) return rc;
  }
  return DPIR_OK;
}

int ensure_graph_staging(dpir_engine *e, int B, int H, int W) {
  size_t need = (size_t)B * 3 * H * W;
  if (need <= e->g_cap) return DPIR_OK;
  if (e->g_x) { cudaFree(e->g_x); cudaFree(e->g_out); cudaFree(e->g_t); }
  // a new staging area invalidates captured graphs
  for (auto &kv : e->plans) {
    if (kv.second->graph_nchw) { cudaGraphExecDestroy(kv.second->graph_nchw); kv.second->graph_nchw = nullptr; }
    if (kv.second->graph_x0) { cudaGraphExecDestroy(kv.second->graph_x0); kv.second->graph_x0 = nullptr; }
  }
  DPIR_CUDA_CHECK(cudaMalloc(&e->g_x, need * 4));
  DPIR_CUDA_CHECK(cudaMalloc(&e->g_out, (size_t)B * e->cfg.out_channels * H * W * 4));
  DPIR_CUDA_CHECK(cudaMalloc(&e->g_t, (size_t)B * 8 + 256));
  e->g_cap = need;
  return DPIR_OK;
}

int build_film_table(dpir_engine *e, cudaStream_t st) {
  const int T = e->cfg.diffusion_steps, ted = e->ted;
  if (!e->film_table) {
    cudaError_t ce = cudaMalloc(&e->film_table, (size_t)T * e->film_rows * sizeof(float));
    if (ce != cudaSuccess) { set_error("FiLM table cudaMalloc failed: %s", cudaGetErrorString(ce)); return DPIR_E_NOMEM; }
  }
  int64_t *tt = nullptr;
  float *hidden = nullptr, *emb = nullptr;
  DPIR_CUDA_CHECK(cudaMalloc(&tt, (size_t)T * sizeof(int64_t)));
  DPIR_CUDA_CHECK(cudaMalloc(&hidden, (size_t)T * ted * sizeof(float)));
  DPIR_CUDA_CHECK(cudaMalloc(&emb, (size_t)T * ted * sizeof(float)));
  std::vector<int64_t> ht(T);
  for (int i = 0; i < T; ++i) ht[i] = i;
  DPIR_CUDA_CHECK(cudaMemcpyAsync(tt, ht.data(), (size_t)T * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  int rc = launch_time_embed(tt, e->ap<float>(e->freqs_off), e->ap<float>(e->te_w0), e->ap<float>(e->te_b0),
                             e->ap<float>(e->te_w2), e->ap<float>(e->te_b2), hidden, emb, T, e->cfg.model_channels, ted, st);
  if (!rc) rc = launch_linear_rows(emb, e->ap<float>(e->film_w), e->ap<float>(e->film_b), e->film_table, T, ted, e->film_rows, 0, 0, st);
  DPIR_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(tt); cudaFree(hidden); cudaFree(emb);
  if (rc) return rc;
  e->film_table_valid = true;
  return DPIR_OK;
}

int forward_impl(dpir_engine *e, const float *x, const int64_t *t, float *out, int B, int H, int W, int mode, void *stream) {
  DPIR_REQUIRE(e && x && t && out, "unet forward: null pointer");
  if (dpir_missing_weights(e) != 0) { set_error("unet forward: %d checkpoint tensors not loaded", dpir_missing_weights(e)); return DPIR_E_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  Plan *plan = nullptr;
  int rc = DPIR_OK;
  if (e->use_film_table && !e->film_table_valid) {
    rc = build_film_table(e, st);
    if (rc) return rc;
  }
  rc = get_plan(e, B, H, W, &plan);
  if (rc) return rc;
  if (!e->graph_mode || e->profiling) {
    if (e->use_film_table) {          // per-timestep tables are indexed with t: work on a clamped copy
      if (e->t_cap < B) {
        if (e->t_buf) { DPIR_CUDA_CHECK(cudaStreamSynchronize(st)); cudaFree(e->t_buf); e->t_buf = nullptr; }
        DPIR_CUDA_CHECK(cudaMalloc(&e->t_buf, (size_t)B * 8 + 256));
        e->t_cap = B;
      }
      rc = launch_clamp_t(t, e->t_buf, B, e->cfg.diffusion_steps, st);
      if (rc) return rc;
      t = e->t_buf;
    }
    return run_plan(e, plan, x, t, out, mode, st);
  }

  // graph mode: fixed staging buffers, one captured graph per (shape, output mode), replayed on the engine's stream
  rc = ensure_graph_staging(e, B, H, W);
  if (rc) return rc;
  if (!e->gstream) {
    DPIR_CUDA_CHECK(cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking));
    DPIR_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming));
    DPIR_CUDA_CHECK(cudaEventCreateWithFlags(&e->ev_out, cudaEventDisableTiming));
  }
  cudaStream_t gs = e->gstream;
  size_t xin = (size_t)B * 3 * H * W * 4;
  size_t xout = (size_t)B * (mode == EPI_X0 ? 3 : e->cfg.out_channels) * H * W * 4;
  cudaGraphExec_t &gexec = mode == EPI_X0 ? plan->graph_x0 : plan->graph_nchw;
  DPIR_CUDA_CHECK(cudaEventRecord(e->ev_in, st));
  DPIR_CUDA_CHECK(cudaStreamWaitEvent(gs, e->ev_in, 0));
  DPIR_CUDA_CHECK(cudaMemcpyAsync(e->g_x, x, xin, cudaMemcpyDeviceToDevice, gs));
  if (e->use_film_table) {
    rc = launch_clamp_t(t, e->g_t, B, e->cfg.diffusion_steps, gs);
    if (rc) return rc;
  } else {
    DPIR_CUDA_CHECK(cudaMemcpyAsync(e->g_t, t, (size_t)B * 8, cudaMemcpyDeviceToDevice, gs));
  }
  if (!gexec) {
    // warm-up run outside capture (sets function attributes, validates the tape), then capture
    rc = run_plan(e, plan, e->g_x, e->g_t, e->g_out, mode, gs);
    if (rc) return rc;
    DPIR_CUDA_CHECK(cudaStreamSynchronize(gs));
    cudaGraph_t graph = nullptr;
    DPIR_CUDA_CHECK(cudaStreamBeginCapture(gs, cudaStreamCaptureModeRelaxed));
    int64_t before = g_launch_count;
    rc = run_plan(e, plan, e->g_x, e->g_t, e->g_out, mode, gs);
    g_launch_count = before;
    cudaError_t ce = cudaStreamEndCapture(gs, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); cudaGetLastError(); return DPIR_E_CUDA; }
    cudaError_t ie = cudaGraphInstantiate(&gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { gexec = nullptr; set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); cudaGetLastError(); return DPIR_E_CUDA; }
  }
  DPIR_CUDA_CHECK(cudaGraphLaunch(gexec, gs));
  // count the kernels the graph replays (same tape)
  for (Op &op : plan->ops) {
    if (op.kind == OP_MEMSET) continue;
    if (op.kind == OP_TIME_EMBED || op.kind == OP_FILM) { g_launch_count += e->use_film_table ? 0 : (op.kind == OP_TIME_EMBED ? 2 : 1); continue; }
    g_launch_count += 1;
  }
  DPIR_CUDA_CHECK(cudaMemcpyAsync(out, e->g_out, xout, cudaMemcpyDeviceToDevice, gs));
  DPIR_CUDA_CHECK(cudaEventRecord(e->ev_out, gs));
  DPIR_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_out, 0));
  return DPIR_OK;
}

}  // namespace

// =============================================================================================== C-ABI
extern "C" {

int dpir_create(const dpir_unet_config *cfg, dpir_engine **out) {
  DPIR_REQUIRE(cfg && out, "dpir_create: null pointer");
  DPIR_REQUIRE(cfg->precision == DPIR_PRECISION_FP32 || cfg->precision == DPIR_PRECISION_FP16 ||
                   cfg->precision == DPIR_PRECISION_FP16S || cfg->precision == DPIR_PRECISION_FP16X3, "dpir_create: bad precision");
  DPIR_REQUIRE(cfg->diffusion_steps >= 1, "dpir_create: diffusion_steps must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("dpir_create: no CUDA device available (this library has no CPU fallback)");
    return DPIR_E_CUDA;
  }
  std::unique_ptr<dpir_engine> e(new dpir_engine());
  e->cfg = *cfg;
  e->half = cfg->precision == DPIR_PRECISION_FP16 || cfg->precision == DPIR_PRECISION_FP16S || cfg->precision == DPIR_PRECISION_FP16X3;
  e->split = cfg->precision == DPIR_PRECISION_FP16X3;
  e->stream_half = cfg->precision == DPIR_PRECISION_FP16S;
  int rc = build_structure(e.get());
  if (rc) return rc;
  rc = build_params(e.get());
  if (rc) return rc;
  cudaError_t ce = cudaMalloc(&e->arena, e->arena_bytes);
  if (ce != cudaSuccess) { set_error("weight arena cudaMalloc(%zu) failed: %s", e->arena_bytes, cudaGetErrorString(ce)); return DPIR_E_NOMEM; }
  DPIR_CUDA_CHECK(cudaMemset(e->arena, 0, e->arena_bytes));
  rc = init_tables(e.get());
  if (rc) { cudaFree(e->arena); return rc; }
  *out = e.release();
  return DPIR_OK;
}

int dpir_destroy(dpir_engine *e) {
  if (!e) return DPIR_OK;
  cudaDeviceSynchronize();
  e->plans.clear();
  if (e->arena) cudaFree(e->arena);
  if (e->film_table) cudaFree(e->film_table);
  if (e->ws) cudaFree(e->ws);
  if (e->g_x) { cudaFree(e->g_x); cudaFree(e->g_out); cudaFree(e->g_t); }
  if (e->t_buf) cudaFree(e->t_buf);
  if (e->gstream) { cudaStreamDestroy(e->gstream); cudaEventDestroy(e->ev_in); cudaEventDestroy(e->ev_out); }
  delete e;
  return DPIR_OK;
}

int dpir_load_weight(dpir_engine *e, const char *name, const float *data, const int64_t *shape, int32_t ndim) {
  DPIR_REQUIRE(e && name && data && shape, "dpir_load_weight: null pointer");
  auto it = e->slots.find(name);
  if (it == e->slots.end()) { set_error("dpir_load_weight: unexpected tensor '%s'", name); return DPIR_E_ARG; }
  ParamSlot &s = it->second;
  bool ok = (int)s.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = s.shape[i] == shape[i];
  // conv1d checkpoints may come as (O,I,1) or (O,I,1,1); 1x1 convs as (O,I,1,1)
  if (!ok) {
    int64_t n_exp = 1, n_got = 1;
    for (auto v : s.shape) n_exp *= v;
    for (int i = 0; i < ndim; ++i) n_got *= shape[i];
    ok = n_exp == n_got && ndim >= 2 && s.shape.size() >= 2 && shape[0] == s.shape[0] && shape[1] == s.shape[1];
  }
  if (!ok) { set_error("dpir_load_weight: shape mismatch for '%s'", name); return DPIR_E_ARG; }
  if (s.kind == PK_COPY || s.kind == PK_FILM_W || s.kind == PK_FILM_B) {
    int64_t n = 1;
    for (auto v : s.shape) n *= v;
    DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.off, data, (size_t)n * 4, cudaMemcpyHostToDevice));
  } else if (s.kind == PK_CONV_STEM) {
    int O = (int)s.shape[0], I = (int)s.shape[1];
    std::vector<float> pk((size_t)9 * I * O);
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I; ++i)
        for (int tp = 0; tp < 9; ++tp) pk[((size_t)tp * I + i) * O + o] = data[((size_t)o * I + i) * 9 + tp];
    DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.off, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
    if (e->stem_tc.cin) {
      std::vector<__half> ph((size_t)9 * O * 64, __float2half(0.f));
      for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
          for (int tp = 0; tp < 9; ++tp) ph[((size_t)tp * O + o) * 64 + i] = __float2half_rn(data[((size_t)o * I + i) * 9 + tp]);
      DPIR_CUDA_CHECK(cudaMemcpy(e->arena + e->stem_tc.w_off, ph.data(), ph.size() * 2, cudaMemcpyHostToDevice));
    }
  } else {  // PK_CONV
    int O = (int)s.shape[0], I = (int)s.shape[1];
    int taps = s.shape.size() == 4 ? (int)(s.shape[2] * s.shape[3]) : 1;
    int ldw = s.ldw;
    if (e->split) {  // [tap][ldw][hi(I) | hi(I) | lo(I)] fp16 of S*w, S a power of two; the epilogue multiplies by 1/S
      float mx = 0.f;
      for (size_t i = 0; i < (size_t)O * I * taps; ++i) mx = std::max(mx, std::fabs(data[i]));
      int ex = 0;
      if (mx > 0.f) std::frexp(mx, &ex);             // mx = m * 2^ex, m in [0.5, 1)
      const float S = std::ldexp(1.0f, 11 - ex);     // S*mx in [1024, 2048): hi and lo stay far from fp16's subnormal range
      std::vector<__half> pk((size_t)taps * ldw * 3 * I, __float2half(0.f));
      for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
          for (int tp = 0; tp < taps; ++tp) {
            const float w = data[((size_t)o * I + i) * taps + tp] * S;
            const __half hi = __float2half_rn(w);
            const __half lo = __float2half_rn(w - __half2float(hi));
            __half *row = pk.data() + ((size_t)tp * ldw + o) * 3 * I;
            row[i] = hi; row[I + i] = hi; row[2 * I + i] = lo;
          }
      DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.off, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
      const float inv = 1.0f / S;
      DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.scale_off, &inv, sizeof(float), cudaMemcpyHostToDevice));
    } else if (e->half) {   // [tap][ldw][I] fp16
      std::vector<__half> pk((size_t)taps * ldw * I, __float2half(0.f));
      for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
          for (int tp = 0; tp < taps; ++tp)
            pk[((size_t)tp * ldw + o) * I + i] = __float2half_rn(data[((size_t)o * I + i) * taps + tp]);
      DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.off, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice));
    } else {         // [tap][I][ldw] fp32
      std::vector<float> pk((size_t)taps * I * ldw, 0.f);
      for (int o = 0; o < O; ++o)
        for (int i = 0; i < I; ++i)
          for (int tp = 0; tp < taps; ++tp) pk[((size_t)tp * I + i) * ldw + o] = data[((size_t)o * I + i) * taps + tp];
      DPIR_CUDA_CHECK(cudaMemcpy(e->arena + s.off, pk.data(), pk.size() * 4, cudaMemcpyHostToDevice));
    }
  }
  s.loaded = true;
  e->film_table_valid = false;
  return DPIR_OK;
}

int dpir_missing_weights(dpir_engine *e) {
  if (!e) return -1;
  int n = 0;
  for (auto &kv : e->slots) n += kv.second.loaded ? 0 : 1;
  return n;
}

int dpir_weight_arena(dpir_engine *e, void **dev_ptr, size_t *bytes) {
  DPIR_REQUIRE(e && dev_ptr && bytes, "dpir_weight_arena: null pointer");
  *dev_ptr = e->arena;
  *bytes = e->arena_bytes;
  return DPIR_OK;
}

int dpir_mark_weights_loaded(dpir_engine *e) {
  DPIR_REQUIRE(e, "dpir_mark_weights_loaded: null engine");
  for (auto &kv : e->slots) kv.second.loaded = true;
  e->film_table_valid = false;
  return DPIR_OK;
}

int dpir_set_tables(dpir_engine *e, const float *freqs, int32_t n_freqs, const float *coef, int32_t T) {
  DPIR_REQUIRE(e, "dpir_set_tables: null engine");
  if (freqs) {
    DPIR_REQUIRE(n_freqs == e->cfg.model_channels / 2, "dpir_set_tables: need model_channels/2 frequencies");
    DPIR_CUDA_CHECK(cudaMemcpy(e->arena + e->freqs_off, freqs, (size_t)n_freqs * 4, cudaMemcpyHostToDevice));
    e->film_table_valid = false;
  }
  if (coef) {
    DPIR_REQUIRE(T == e->cfg.diffusion_steps, "dpir_set_tables: coefficient table must have diffusion_steps rows");
    DPIR_CUDA_CHECK(cudaMemcpy(e->arena + e->coef_off, coef, (size_t)T * 2 * 4, cudaMemcpyHostToDevice));
  }
  return DPIR_OK;
}

int dpir_unet_forward(dpir_engine *e, const float *x, const int64_t *t, float *out, int32_t B, int32_t H, int32_t W, void *stream) {
  return forward_impl(e, x, t, out, B, H, W, EPI_NCHW, stream);
}

int dpir_unet_x0(dpir_engine *e, const float *x, const int64_t *t, float *x0, int32_t B, int32_t H, int32_t W, void *stream) {
  DPIR_REQUIRE(e && e->cfg.in_channels == 3, "dpir_unet_x0: needs a 3-channel model");
  return forward_impl(e, x, t, x0, B, H, W, EPI_X0, stream);
}

int64_t dpir_workspace_bytes(dpir_engine *e) { return e ? (int64_t)e->ws_bytes : 0; }

int dpir_set_film_table(dpir_engine *e, int32_t enable) {
  DPIR_REQUIRE(e, "dpir_set_film_table: null engine");
  if ((e->use_film_table != 0) != (enable != 0)) {
    // captured graphs bake the choice in: drop them
    for (auto &kv : e->plans) {
      if (kv.second->graph_nchw) { cudaGraphExecDestroy(kv.second->graph_nchw); kv.second->graph_nchw = nullptr; }
      if (kv.second->graph_x0) { cudaGraphExecDestroy(kv.second->graph_x0); kv.second->graph_x0 = nullptr; }
    }
  }
  e->use_film_table = enable ? 1 : 0;
  return DPIR_OK;
}

int dpir_set_graph_mode(dpir_engine *e, int32_t enable) {
  DPIR_REQUIRE(e, "dpir_set_graph_mode: null engine");
  e->graph_mode = enable ? 1 : 0;
  return DPIR_OK;
}

int dpir_profile_begin(dpir_engine *e) {
  DPIR_REQUIRE(e, "dpir_profile_begin: null engine");
  for (auto ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear();
  e->prof_op.clear();
  e->prof_plan = nullptr;
  e->profiling = true;
  return DPIR_OK;
}

int dpir_profile_end(dpir_engine *e, int32_t *kind, int32_t *dims, float *ms, int32_t cap) {
  DPIR_REQUIRE(e, "dpir_profile_end: null engine");
  e->profiling = false;
  DPIR_CUDA_CHECK(cudaDeviceSynchronize());
  int n = (int)e->prof_op.size();
  int nops = e->prof_plan ? (int)e->prof_plan->ops.size() : 0;
  // accumulate over repeated forwards of the same plan: one row per tape op
  std::vector<float> acc(nops, 0.f);
  for (int i = 0; i < n; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, e->prof_events[2 * i], e->prof_events[2 * i + 1]);
    if (e->prof_op[i] < nops) acc[e->prof_op[i]] += t;
  }
  for (auto ev : e->prof_events) cudaEventDestroy(ev);
  e->prof_events.clear();
  e->prof_op.clear();
  int rows = nops < cap ? nops : cap;
  for (int i = 0; i < rows && kind && dims && ms; ++i) {
    const Op &op = e->prof_plan->ops[i];
    kind[i] = (int)op.kind;
    int H = op.H, W = op.W, Cin = op.Cin_alg ? op.Cin_alg : op.Cin, Cout = op.Cout;
    if (op.kind == OP_GN_APPLY) { H = op.gn.H; W = op.gn.W; Cin = op.gn.C1 + op.gn.C2; Cout = op.gn.resample; }
    if (op.kind == OP_GN_STATS) { H = op.HW; W = 1; Cin = op.C1 + op.C2; Cout = 0; }
    if (op.kind == OP_ATTN || op.kind == OP_SPLIT_CAST) { H = op.HW; W = 1; Cout = op.heads; }
    dims[6 * i] = H; dims[6 * i + 1] = W; dims[6 * i + 2] = Cin; dims[6 * i + 3] = Cout;
    dims[6 * i + 4] = op.kind == OP_CONV ? op.ksize + (op.gf.coef ? 10 : 0) : 0;   // 13 = 3x3 with the GroupNorm+SiLU fused on its operand
    dims[6 * i + 5] = op.kind == OP_CONV ? op.Cin2 : 0;
    ms[i] = acc[i];
  }
  e->prof_plan = nullptr;
  return nops;
}

// ---- per-kernel entry points for tests / benchmarks
int dpir_conv_tc(const void *act, const void *w, const float *bias, const float *residual, float *out, int32_t B, int32_t H,
                 int32_t W, int32_t Cin, int32_t Cout, int32_t CoutPad, int32_t ksize, void *stream) {
  DPIR_REQUIRE(act && w && bias && out, "dpir_conv_tc: null pointer");
  ConvTcPlan *p = nullptr;
  int rc = conv_tc_plan_create(act, w, B, H, W, Cin, CoutPad, ksize, &p);
  if (rc) return rc;
  ConvEpi epi{};
  epi.mode = EPI_NHWC; epi.bias = bias; epi.residual = residual; epi.out = out;
  rc = launch_conv_tc(p, epi, Cout, (cudaStream_t)stream);
  conv_tc_plan_destroy(p);   // the tensor maps were copied into the launch parameters
  return rc;
}

/* test entry for the fp16-stream form of the same convolution: fp16 residual (or NULL), fp16 NHWC output (this is the
 * combination the swap-AB persistent kernel serves with its TMA-store epilogue) */
int dpir_conv_tc_f16io(const void *act, const void *w, const float *bias, const void *residual_half, void *out_half, int32_t B,
                       int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t CoutPad, int32_t ksize, void *stream) {
  DPIR_REQUIRE(act && w && bias && out_half, "dpir_conv_tc_f16io: null pointer");
  ConvTcPlan *p = nullptr;
  int rc = conv_tc_plan_create(act, w, B, H, W, Cin, CoutPad, ksize, &p);
  if (rc) return rc;
  ConvEpi epi{};
  epi.mode = EPI_NHWC; epi.bias = bias; epi.residual = reinterpret_cast<const float *>(residual_half);
  epi.res_half = residual_half ? 1 : 0; epi.out = reinterpret_cast<float *>(out_half); epi.out_half = 1;
  rc = conv_tc_plan_set_io(p, out_half, 1, residual_half, 1, Cout);
  if (!rc) rc = launch_conv_tc(p, epi, Cout, (cudaStream_t)stream);
  conv_tc_plan_destroy(p);
  return rc;
}

/* 1 (default): ResBlocks whose 3x3 convs run on the persistent swap-AB kernel apply GroupNorm(+FiLM)+SiLU inside the conv
 * (no k_gn_apply pass, fp16s pipeline only), fp32 operand arithmetic; 2: the same with packed-half arithmetic; 0: every norm
 * is a separate pass.  Cached plans of `e` are dropped. */
int dpir_set_gn_fusion(dpir_engine *e, int32_t enable) {
  DPIR_REQUIRE(enable >= 0 && enable <= 4,
               "dpir_set_gn_fusion: 0 = off, 1 | 3 = on (activate-once producers), 2 = per-tap packed half, 4 = per-tap fp32");
  const int math = enable == 2 ? 2 : (enable == 4 ? 1 : 3);
  if (e) {                       // per engine: other engines of the process keep their setting
    e->gn_fuse = enable ? 1 : 0;
    if (enable) e->gn_math = math;
    cudaDeviceSynchronize();
    e->plans.clear();
  } else if (enable) {           // no engine: default of the per-kernel test entry dpir_conv_gnfuse_f16io
    conv_tc_set_gnfuse_math(math);
  }
  return DPIR_OK;
}

/* test entries for the fused-operand (v4) convolution: dpir_gn_coef folds quad statistics + affine (+ per-sample FiLM rows
 * film[b*film_stride + {c, C+c}]) into the [B][2][C] coefficient table; dpir_conv_gnfuse_f16io runs
 * out = conv3x3(SiLU(2*(x*hs+ho))) (+ residual) with x = (src1 | src2) raw fp16 NHWC, fp16 NHWC output */
int dpir_gn_coef(const double *q1, const double *q2, int32_t C1, int32_t C2, const float *gamma, const float *beta, const float *film,
                 int32_t film_stride, int32_t B, int32_t HW, float *coef, void *stream) {
  DPIR_REQUIRE(q1 && gamma && beta && coef, "dpir_gn_coef: null pointer");
  GnCoefArgs a{};
  a.C1 = C1; a.C2 = C2; a.q1 = q1; a.q2 = q2; a.gamma = gamma; a.beta = beta; a.film = film; a.film_stride = film_stride;
  a.tsteps = nullptr; a.B = B; a.HW = HW; a.coef = coef;
  return launch_gn_coef(a, (cudaStream_t)stream);
}

int dpir_conv_gnfuse_f16io(const void *src1, int32_t C1, const void *src2, int32_t C2, const float *coef, const void *w,
                           const float *bias, const void *residual_half, void *out_half, int32_t B, int32_t H, int32_t W,
                           int32_t Cout, int32_t CoutPad, void *stream) {
  DPIR_REQUIRE(src1 && coef && w && bias && out_half, "dpir_conv_gnfuse_f16io: null pointer");
  ConvTcPlan *p = nullptr;
  int rc = conv_tc_plan_create(src1, w, B, H, W, C1 + C2, CoutPad, 3, &p);
  if (rc) return rc;
  if (CoutPad == 16) {      // narrow-output kernel (the UNet's output layer): one source, no residual, fp32 NHWC output
    ConvEpi epi{};
    epi.mode = EPI_NHWC; epi.bias = bias; epi.out = reinterpret_cast<float *>(out_half);
    GnFuse gf;
    gf.src1 = (const __half *)src1; gf.C1 = C1; gf.src2 = (const __half *)src2; gf.C2 = C2; gf.coef = coef;
    if (residual_half) { set_error("dpir_conv_gnfuse_f16io: the narrow-output form (CoutPad 16) takes no residual"); rc = DPIR_E_ARG; }
    else rc = conv_tc_plan_set_gn_fuse(p, gf);
    if (!rc) rc = launch_conv_tc(p, epi, Cout, (cudaStream_t)stream);
    conv_tc_plan_destroy(p);
    return rc;
  }
  ConvEpi epi{};
  epi.mode = EPI_NHWC; epi.bias = bias; epi.residual = reinterpret_cast<const float *>(residual_half);
  epi.res_half = residual_half ? 1 : 0; epi.out = reinterpret_cast<float *>(out_half); epi.out_half = 1;
  rc = conv_tc_plan_set_io(p, out_half, 1, residual_half, 1, Cout);
  GnFuse gf;
  gf.src1 = (const __half *)src1; gf.C1 = C1; gf.src2 = (const __half *)src2; gf.C2 = C2; gf.coef = coef;
  if (!rc) rc = conv_tc_plan_set_gn_fuse(p, gf);
  if (!rc) rc = launch_conv_tc(p, epi, Cout, (cudaStream_t)stream);
  conv_tc_plan_destroy(p);
  return rc;
}

/* QKVAttentionLegacy / QKVAttention (unet.py:337-354, 371-389) on its own: qkv dev fp32 [B][T][3C] (NHWC output of the qkv 1x1
 * convolution; legacy channel order [head][q,k,v][64], new order [q,k,v][head][64]), out dev [B][T][C] fp16 (out_half: the
 * tcgen05 kernel) or fp32 (the fp32 CUDA-core kernel).  Any T >= 1. */
int dpir_attention(const float *qkv, void *out, int32_t out_half, int32_t B, int32_t T, int32_t C, int32_t heads, int32_t new_order,
                   void *stream) {
  DPIR_REQUIRE(qkv && out, "dpir_attention: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (!out_half) return launch_attention(qkv, out, 0, B, T, C, heads, new_order, st);
  // tensor-core path: the engine's qkv convolution writes fp16; this test entry rounds its fp32 input the same way (stream-ordered
  // temporary) and runs the tcgen05 kernel on it
  DPIR_REQUIRE(C == heads * 64 && T >= 1 && B >= 1, "dpir_attention: head dim must be 64");
  __half *tmp = nullptr;
  const int64_t n = (int64_t)B * T * 3 * C;
  DPIR_CUDA_CHECK(cudaMallocAsync(&tmp, (size_t)n * sizeof(__half), st));
  int rc = launch_cast_half(qkv, tmp, n, st);
  AttnTcPlan *ap = nullptr;
  if (!rc) rc = attn_tc_plan_create(tmp, B, T, C, &ap);
  if (!rc) rc = launch_attention_tc(ap, out, 1, heads, new_order, st);
  if (ap) attn_tc_plan_destroy(ap);          // the tensor map was copied into the launch parameters
  cudaFreeAsync(tmp, st);
  return rc;
}

int dpir_conv_tc_variant(int32_t variant) {
  DPIR_REQUIRE(variant >= 0 && variant <= 2, "dpir_conv_tc_variant: 0 = auto, 1 = v1, 2 = v2");
  conv_tc_force_v1(variant);
  return DPIR_OK;
}

int dpir_conv_simt(const float *act, const float *w, const float *bias, const float *residual, float *out, int32_t B, int32_t H,
                   int32_t W, int32_t Cin, int32_t Cout, int32_t ksize, void *stream) {
  DPIR_REQUIRE(act && w && bias && out, "dpir_conv_simt: null pointer");
  ConvEpi epi{};
  epi.mode = EPI_NHWC; epi.bias = bias; epi.residual = residual; epi.out = out;
  return launch_conv_simt(act, w, (Cout + 3) / 4 * 4, epi, B, H, W, Cin, Cout, ksize, (cudaStream_t)stream);
}

}  // extern "C"
