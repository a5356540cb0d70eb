#include "caffe/dc_engine.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <sstream>

#include "caffe/layers/dc_layers.hpp"
#include "deepcut_b200.h"

namespace caffe {

typedef BaseConvolutionLayer<float> ConvBase;

// A value flowing through the net.  In-place layers create a new Tensor on the same blob; Split
// tops alias their bottom's Tensor.
struct FusedPlan::Tensor {
  enum Kind { kBlobF32, kSplit, kF32Rows, kRaw, kVirtual } kind = kVirtual;
  size_t raw_bytes = 0;               // kRaw: scratch of this many bytes
  int id = 0;
  int n = 0, c = 0, h = 0, w = 0;
  int ld = 0;                         // kF32Rows: row stride (floats); rows = n*h*w
  int blob = -1;                      // Net blob that names this value (-1: internal)
  int producer_layer = -1;            // -1: net input
  std::vector<int> consumers;         // layer ids reading it (Split layers excluded)
  int def_step = -1, last_step = -1;  // liveness in step indices
  int alloc_step = -1, free_step = -1;  // arena residency (liveness widened to whole segments for tensors crossing a chunked one)
  int chunk_n = 0;                    // > 0: lives entirely inside a chunked segment; the arena holds this many images of it
  bool gave_memory = false;           // its storage was taken over in place by the step that read it last (block output over shortcut)
  size_t bytes = 0, offset = 0;       // arena placement
  void* ptr = nullptr;                // resolved device address (arena tensors)
  size_t elems() const { return static_cast<size_t>(n) * c * h * w; }
};

struct FusedPlan::Step {
  enum Type { kConv1, kConvBN, kSubsample, kMaxPool, kHeadGemm, kHeadFinish, kToBlob } type = kConvBN;
  std::string name;
  Tensor* in = nullptr;
  Tensor* in2 = nullptr;              // residual (kConvBN) / skip rows (kHeadFinish)
  Tensor* out = nullptr;
  // convolution
  int conv_layer = -1, bn_layer = -1, scale_layer = -1;
  bool relu = false;
  int kh = 1, kw = 1, pad = 0, dil = 1, stride = 1, cout = 0;
  bool deconv_rows = false;           // kHeadGemm over deconv weights
  std::vector<int> merged_layers;     // kHeadGemm: layers whose weights are concatenated
  void* w_dev = nullptr;
  float* scale_dev = nullptr;
  float* shift_dev = nullptr;
  // head finish
  int col_off = 0, skip_off = 0, sigmoid = 0, out_blob = -1;
  Tensor* col = nullptr;
  Tensor* ws = nullptr;               // kConv1 (tensor-core stem): space-to-depth scratch
  bool stem_tc = false;
  // pooling
  int pool_k = 3, pool_s = 2;
};

FusedPlan::~FusedPlan() {
  for (Tensor* t : tensors_) delete t;
  for (Step* s : steps_) delete s;
  for (void* e : events_) dc_event_destroy(e);
  if (graph_) dc_graph_destroy(graph_);
  if (arena_) dc_free(arena_);
}

namespace {

template <class L>
L* As(Layer<float>* l) { return dynamic_cast<L*>(l); }

struct Matcher {
  Net<float>& net;
  std::vector<FusedPlan::Tensor*>& tensors;
  std::vector<FusedPlan::Step*>& steps;
  std::vector<int> cur;                       // blob index -> tensor id currently held
  std::vector<std::vector<int> > bot_t, top_t;   // per layer tensor ids
  std::vector<char> done;
  std::string why;

  Matcher(Net<float>& n, std::vector<FusedPlan::Tensor*>& t, std::vector<FusedPlan::Step*>& s) : net(n), tensors(t), steps(s) {}

  FusedPlan::Tensor* NewTensor(int blob, int producer) {
    FusedPlan::Tensor* t = new FusedPlan::Tensor();
    t->id = static_cast<int>(tensors.size());
    t->blob = blob;
    t->producer_layer = producer;
    if (blob >= 0) {
      const Blob<float>& b = *net.blobs()[blob];
      if (b.num_axes() == 4) { t->n = b.shape(0); t->c = b.shape(1); t->h = b.shape(2); t->w = b.shape(3); }
    }
    tensors.push_back(t);
    return t;
  }
  FusedPlan::Tensor* NewInternal(FusedPlan::Tensor::Kind k, int n, int c, int h, int w) {
    FusedPlan::Tensor* t = NewTensor(-1, -2);
    t->kind = k; t->n = n; t->c = c; t->h = h; t->w = w;
    return t;
  }
  bool Fail(const std::string& m) { why = m; return false; }

  void BuildDataflow() {
    const int nl = static_cast<int>(net.layers().size());
    cur.assign(net.blobs().size(), -1);
    bot_t.resize(nl);
    top_t.resize(nl);
    done.assign(nl, 0);
    for (int idx : net.input_blob_indices()) {
      FusedPlan::Tensor* t = NewTensor(idx, -1);
      t->kind = FusedPlan::Tensor::kBlobF32;
      cur[idx] = t->id;
    }
    for (int i = 0; i < nl; ++i) {
      const bool is_split = std::string(net.layers()[i]->type()) == "Split";
      for (int b : net.bottom_ids(i)) {
        bot_t[i].push_back(cur[b]);
        if (!is_split) tensors[cur[b]]->consumers.push_back(i);
      }
      for (int b : net.top_ids(i)) {
        if (is_split) { cur[b] = bot_t[i][0]; top_t[i].push_back(cur[b]); continue; }
        FusedPlan::Tensor* t = NewTensor(b, i);
        cur[b] = t->id;
        top_t[i].push_back(t->id);
      }
    }
    // values still held by output blobs are consumed by the caller
    for (int idx : net.output_blob_indices()) tensors[cur[idx]]->consumers.push_back(-1);
  }

  // the single consumer of tensor t if it is layer type `type` operating in place, else -1
  int InPlaceNext(int t, const char* type) {
    const std::vector<int>& c = tensors[t]->consumers;
    if (c.size() != 1 || c[0] < 0) return -1;
    const int j = c[0];
    if (std::string(net.layers()[j]->type()) != type) return -1;
    if (net.bottom_ids(j).size() != 1 || net.top_ids(j).size() != 1 || net.bottom_ids(j)[0] != net.top_ids(j)[0]) return -1;
    return j;
  }

  FusedPlan::Step* AddStep(FusedPlan::Step::Type type, const std::string& name) {
    FusedPlan::Step* s = new FusedPlan::Step();
    s->type = type;
    s->name = name;
    steps.push_back(s);
    return s;
  }

  std::map<std::pair<int, int>, FusedPlan::Tensor*> subsampled;   // (tensor, stride) -> gathered tensor

  bool MatchConv(int i) {
    ConvBase* conv = As<ConvBase>(net.layers()[i].get());
    const std::string lname = net.layer_names()[i];
    if (net.bottom_ids(i).size() != 1) return Fail("convolution " + lname + " has several bottoms");
    FusedPlan::Tensor* in = tensors[bot_t[i][0]];
    int t = top_t[i][0];
    int bn = InPlaceNext(t, "BatchNorm");
    if (bn >= 0) t = top_t[bn][0];
    int sc = InPlaceNext(t, "Scale");
    if (sc >= 0) {
      // only the per-channel form folds into the epilogue: gamma (and beta) of shape {Cout} applied along axis 1
      Layer<float>* sl = net.layers()[sc].get();
      const ScaleParameter& sp = sl->layer_param().scale_param();
      const bool per_channel = sl->blobs().size() >= 1 && sl->blobs()[0]->count() == conv->num_output() && sl->blobs()[0]->num_axes() == 1 &&
                               sp.axis() == 1 && (sl->blobs().size() < 2 || sl->blobs()[1]->count() == conv->num_output());
      if (per_channel) t = top_t[sc][0];
      else sc = -1;
    }
    int rl = InPlaceNext(t, "ReLU");
    if (rl >= 0) {
      if (net.layers()[rl]->layer_param().relu_param().negative_slope() != 0.f) rl = -1;   // leaky: leave to the per-layer path
      else t = top_t[rl][0];
    }
    if (conv->kernel_h() != conv->kernel_w() || conv->pad_h() != conv->pad_w() || conv->stride_h() != conv->stride_w() ||
        conv->dilation_h() != conv->dilation_w())
      return Fail("convolution " + lname + " is not square");
    const int k = conv->kernel_h(), s = conv->stride_h(), p = conv->pad_h(), d = conv->dilation_h();
    FusedPlan::Tensor* out = tensors[t];
    done[i] = 1;
    if (bn >= 0) done[bn] = 1;
    if (sc >= 0) done[sc] = 1;
    if (rl >= 0) done[rl] = 1;
    // intermediate in-place values (conv raw output, post-BN, ...) never exist
    for (int v = top_t[i][0]; v != t;) {
      tensors[v]->kind = FusedPlan::Tensor::kVirtual;
      v = top_t[tensors[v]->consumers[0]][0];
    }
    if (in->kind == FusedPlan::Tensor::kBlobF32) {
      if (!(in->c == 3 && k == 7 && s == 2 && p == 3 && d == 1 && conv->num_output() == 64 && rl >= 0 && !conv->bias_term()))
        return Fail("convolution " + lname + " reads an fp32 blob but is not the 7x7/2 3->64 stem");
      FusedPlan::Step* st = AddStep(FusedPlan::Step::kConv1, lname);
      st->in = in; st->out = out; st->conv_layer = i; st->bn_layer = bn; st->scale_layer = sc; st->relu = true;
      st->cout = 64;
      out->kind = FusedPlan::Tensor::kSplit;
      const char* e = getenv("DC_STEM_TC");
      st->stem_tc = !(e && e[0] == '0');
      if (st->stem_tc) {
        st->ws = NewInternal(FusedPlan::Tensor::kRaw, in->n, 16, (in->h + 1) / 2, (in->w + 1) / 2 + 3);
        st->ws->raw_bytes = dc_conv1_tc_workspace_bytes(in->n, in->h, in->w);
      }
      return true;
    }
    if (in->kind != FusedPlan::Tensor::kSplit) return Fail("convolution " + lname + " input is not a split-NHWC activation");
    if (in->c % 64 != 0) return Fail("convolution " + lname + ": cin not a multiple of 64");
    if (conv->num_output() % 32 != 0) return Fail("convolution " + lname + ": cout not a multiple of 32");
    if (k * k > 9) return Fail("convolution " + lname + ": more than 9 taps");
    if (conv->bias_term() && (bn >= 0 || sc >= 0)) return Fail("convolution " + lname + ": bias followed by BatchNorm");
    int conv_stride = 1;
    if (s != 1) {
      if (!(k == 1 && p == 0)) return Fail("convolution " + lname + ": stride > 1 only for 1x1 pad 0");
      // default: the conv reads every s-th pixel itself (TMA traversal stride); DC_TMA_STRIDE=0 gathers first
      static const bool tma_stride = [] { const char* e = getenv("DC_TMA_STRIDE"); return !(e && e[0] == '0'); }();
      const std::pair<int, int> key(in->id, s);
      if (tma_stride && s <= 8) {
        conv_stride = s;
      } else if (!subsampled.count(key)) {
        FusedPlan::Tensor* g = NewInternal(FusedPlan::Tensor::kSplit, in->n, in->c, (in->h - 1) / s + 1, (in->w - 1) / s + 1);
        FusedPlan::Step* ss = AddStep(FusedPlan::Step::kSubsample, lname + "/gather");
        ss->in = in; ss->out = g; ss->stride = s;
        subsampled[key] = g;
      }
      if (conv_stride == 1) in = subsampled[key];
    }
    FusedPlan::Step* st = AddStep(FusedPlan::Step::kConvBN, lname);
    st->stride = conv_stride;
    st->in = in; st->out = out; st->conv_layer = i; st->bn_layer = bn; st->scale_layer = sc; st->relu = rl >= 0;
    st->kh = st->kw = k; st->pad = p; st->dil = d; st->cout = conv->num_output();
    out->kind = FusedPlan::Tensor::kSplit;
    return true;
  }

  int StepProducing(FusedPlan::Tensor* t) {
    for (int s = static_cast<int>(steps.size()) - 1; s >= 0; --s)
      if (steps[s]->out == t) return s;
    return -1;
  }

  bool MatchEltwise(int i) {
    EltwiseLayer<float>* e = As<EltwiseLayer<float> >(net.layers()[i].get());
    const std::string lname = net.layer_names()[i];
    if (bot_t[i].size() != 2 || e->coeffs()[0] != 1.f || e->coeffs()[1] != 1.f) return Fail("eltwise " + lname + " is not a plain 2-input sum");
    // which bottom is the branch (a ConvBN output without ReLU, read only here)?
    for (int side = 1; side >= 0; --side) {
      FusedPlan::Tensor* branch = tensors[bot_t[i][side]];
      FusedPlan::Tensor* other = tensors[bot_t[i][1 - side]];
      const int sb = StepProducing(branch);
      if (sb < 0 || steps[sb]->type != FusedPlan::Step::kConvBN || steps[sb]->relu || steps[sb]->in2 != nullptr) continue;
      if (branch->consumers.size() != 1 || other->kind != FusedPlan::Tensor::kSplit) continue;
      const int so = StepProducing(other);
      if (so >= sb) continue;                 // shortcut must exist before the branch conv runs
      int t = top_t[i][0];
      const int rl = InPlaceNext(t, "ReLU");
      if (rl >= 0 && net.layers()[rl]->layer_param().relu_param().negative_slope() == 0.f) {
        tensors[t]->kind = FusedPlan::Tensor::kVirtual;
        t = top_t[rl][0];
        done[rl] = 1;
        steps[sb]->relu = true;
      }
      branch->kind = FusedPlan::Tensor::kVirtual;
      steps[sb]->in2 = other;
      steps[sb]->out = tensors[t];
      steps[sb]->name += "+" + lname;
      tensors[t]->kind = FusedPlan::Tensor::kSplit;
      done[i] = 1;
      return true;
    }
    return Fail("eltwise " + lname + " does not close a residual branch the conv epilogue can absorb");
  }

  bool MatchPool(int i) {
    PoolingLayer<float>* pl = As<PoolingLayer<float> >(net.layers()[i].get());
    const std::string lname = net.layer_names()[i];
    FusedPlan::Tensor* in = tensors[bot_t[i][0]];
    if (in->kind != FusedPlan::Tensor::kSplit) return Fail("pooling " + lname + " input is not a split-NHWC activation");
    if (pl->kernel_h() != pl->kernel_w() || pl->stride_h() != pl->stride_w() || pl->pad_h() != 0 || pl->pad_w() != 0)
      return Fail("pooling " + lname + ": only square, unpadded windows");
    FusedPlan::Step* st = AddStep(FusedPlan::Step::kMaxPool, lname);
    st->in = in; st->out = tensors[top_t[i][0]];
    st->pool_k = pl->kernel_h(); st->pool_s = pl->stride_h();
    st->out->kind = FusedPlan::Tensor::kSplit;
    done[i] = 1;
    return true;
  }

  // All heads  Deconvolution(X5) -> Crop(., skip) ; skip = Convolution1x1(X3) ; Eltwise(skip, crop) [; Sigmoid]
  bool MatchHeads(int first) {
    struct Head { int deconv, skipconv, crop, elt, sig; };
    std::vector<Head> heads;
    FusedPlan::Tensor* x5 = tensors[bot_t[first][0]];
    FusedPlan::Tensor* x3 = nullptr;
    const int nl = static_cast<int>(net.layers().size());
    for (int d = first; d < nl; ++d) {
      if (done[d] || std::string(net.layers()[d]->type()) != "Deconvolution" || tensors[bot_t[d][0]] != x5) continue;
      ConvBase* dc = As<ConvBase>(net.layers()[d].get());
      const std::string dn = net.layer_names()[d];
      if (!(dc->kernel_h() == 3 && dc->kernel_w() == 3 && dc->stride_h() == 2 && dc->stride_w() == 2 && dc->pad_h() == 0 && dc->pad_w() == 0 &&
            dc->dilation_h() == 1 && dc->dilation_w() == 1))
        return Fail("deconvolution " + dn + " is not 3x3 stride 2 pad 0");
      FusedPlan::Tensor* up = tensors[top_t[d][0]];
      if (up->consumers.size() != 1 || up->consumers[0] < 0 || std::string(net.layers()[up->consumers[0]]->type()) != "Crop")
        return Fail("deconvolution " + dn + " is not followed by Crop");
      Head hd;
      hd.deconv = d;
      hd.crop = up->consumers[0];
      CropLayer<float>* cl = As<CropLayer<float> >(net.layers()[hd.crop].get());
      if (cl->crop_h() != 0 || cl->crop_w() != 0 || bot_t[hd.crop][0] != up->id) return Fail("crop after " + dn + " has non-zero offsets");
      FusedPlan::Tensor* skip = tensors[bot_t[hd.crop][1]];
      hd.skipconv = skip->producer_layer;
      if (hd.skipconv < 0 || std::string(net.layers()[hd.skipconv]->type()) != "Convolution") return Fail("crop reference of " + dn + " is not a convolution output");
      ConvBase* sk = As<ConvBase>(net.layers()[hd.skipconv].get());
      if (!(sk->kernel_h() == 1 && sk->kernel_w() == 1 && sk->stride_h() == 1 && sk->pad_h() == 0 && sk->num_output() == dc->num_output()))
        return Fail("skip head of " + dn + " is not a 1x1 convolution with matching outputs");
      FusedPlan::Tensor* sin = tensors[bot_t[hd.skipconv][0]];
      if (x3 == nullptr) x3 = sin;
      if (sin != x3 || sin->kind != FusedPlan::Tensor::kSplit) return Fail("heads do not share one skip input");
      FusedPlan::Tensor* cropped = tensors[top_t[hd.crop][0]];
      // skip is read by the crop (shape only) and the eltwise; cropped only by the eltwise
      if (cropped->consumers.size() != 1 || cropped->consumers[0] < 0 || std::string(net.layers()[cropped->consumers[0]]->type()) != "Eltwise")
        return Fail("crop of " + dn + " does not feed an Eltwise");
      hd.elt = cropped->consumers[0];
      EltwiseLayer<float>* e = As<EltwiseLayer<float> >(net.layers()[hd.elt].get());
      if (bot_t[hd.elt].size() != 2 || e->coeffs()[0] != 1.f || e->coeffs()[1] != 1.f) return Fail("head eltwise is not a plain sum");
      const int o0 = bot_t[hd.elt][0], o1 = bot_t[hd.elt][1];
      if (!((o0 == skip->id && o1 == cropped->id) || (o1 == skip->id && o0 == cropped->id))) return Fail("head eltwise does not add skip and cropped deconv");
      for (int c : skip->consumers) if (c != hd.crop && c != hd.elt) return Fail("skip head output has other readers");
      hd.sig = -1;
      FusedPlan::Tensor* sum = tensors[top_t[hd.elt][0]];
      if (sum->consumers.size() == 1 && sum->consumers[0] >= 0 && std::string(net.layers()[sum->consumers[0]]->type()) == "Sigmoid") hd.sig = sum->consumers[0];
      heads.push_back(hd);
    }
    if (heads.empty()) return Fail("no head matched");
    // heads whose final blob the caller declared unread (Net::set_skipped_outputs) are matched -- their layers count as done,
    // their values as fused away -- but take no rows of the merged GEMMs and get no finishing step
    std::vector<Head> skipped;
    std::string skip_tag;
    if (!net.skipped_outputs().empty()) {
      std::vector<Head> kept;
      for (const Head& h : heads) {
        const FusedPlan::Tensor* o = tensors[top_t[h.sig >= 0 ? h.sig : h.elt][0]];
        const bool skip = o->blob >= 0 && net.skipped_outputs().count(net.blob_names()[o->blob]) != 0;
        (skip ? skipped : kept).push_back(h);
        if (skip) skip_tag += "-" + net.blob_names()[o->blob];
      }
      if (kept.empty()) return Fail("every head output is in the skipped set");
      heads.swap(kept);
    }
    for (const Head& h : skipped) {
      for (int l : {h.deconv, h.skipconv, h.crop, h.elt}) { tensors[top_t[l][0]]->kind = FusedPlan::Tensor::kVirtual; done[l] = 1; }
      if (h.sig >= 0) { tensors[top_t[h.sig][0]]->kind = FusedPlan::Tensor::kVirtual; done[h.sig] = 1; }
    }
    int ctot = 0;
    for (const Head& h : heads) ctot += As<ConvBase>(net.layers()[h.deconv].get())->num_output();
    // what the merged channel-major GEMMs (dc_conv_forward, out_f32_rows = 2) need: 128-row weight tiles, i.e. more than 64
    // merged outputs, and inputs in whole 64-channel K chunks.  A net trimmed to e.g. the part + locref heads (14 + 28 rows)
    // runs layer by layer instead of failing inside the plan.
    // The merged channel-major GEMMs (dc_conv_forward, out_f32_rows = 2) run 128-row weight tiles: a net trimmed to e.g. the
    // part + locref heads (14 + 28 outputs; callers that never read next_pred, estimate_pose.py:231) gets its skip matrix
    // zero-padded to a full tile instead of being refused.
    const int skip_rows = std::max(ctot, 65);
    if (x5->kind != FusedPlan::Tensor::kSplit || x5->c % 64 != 0 || x3->c % 64 != 0)
      return Fail("head inputs must be split activations with a multiple of 64 channels");
    const int h5 = x5->h, w5 = x5->w, h3 = x3->h, w3 = x3->w;
    if (!(2 * h5 + 1 > h3 && 2 * w5 + 1 > w3 && h3 <= 2 * h5 + 1)) return Fail("head geometry: crop larger than the deconvolution output");
    // merged deconv GEMM -> col rows, merged 1x1 GEMM -> skip rows
    // channel-major fp32 results: rows = packed weight rows, ld = pixels rounded up to 32
    auto round32 = [](long long v) { return static_cast<int>((v + 31) / 32 * 32); };
    FusedPlan::Tensor* col = NewInternal(FusedPlan::Tensor::kF32Rows, 1, 1, dc_packed_rows(ctot * 9), 1);
    col->ld = round32(static_cast<long long>(x5->n) * h5 * w5);
    // (the names carry the skipped set: the packed-weight cache is keyed by step type + name)
    FusedPlan::Step* g1 = AddStep(FusedPlan::Step::kHeadGemm, "heads/deconv_gemm" + skip_tag);
    g1->in = x5; g1->out = col; g1->deconv_rows = true; g1->cout = ctot * 9;
    FusedPlan::Tensor* srows = NewInternal(FusedPlan::Tensor::kF32Rows, 1, 1, dc_packed_rows(skip_rows), 1);
    srows->ld = round32(static_cast<long long>(x3->n) * h3 * w3);
    FusedPlan::Step* g2 = AddStep(FusedPlan::Step::kHeadGemm, "heads/skip_gemm" + skip_tag);
    g2->in = x3; g2->out = srows; g2->deconv_rows = false; g2->cout = skip_rows;
    int off = 0;
    for (const Head& h : heads) {
      g1->merged_layers.push_back(h.deconv);
      g2->merged_layers.push_back(h.skipconv);
      const int co = As<ConvBase>(net.layers()[h.deconv].get())->num_output();
      const int out_layer = h.sig >= 0 ? h.sig : h.elt;
      FusedPlan::Tensor* out = tensors[top_t[out_layer][0]];
      FusedPlan::Step* f = AddStep(FusedPlan::Step::kHeadFinish, net.layer_names()[out_layer]);
      f->col = col; f->in2 = srows; f->in = x5; f->out = out; f->col_off = off * 9; f->skip_off = off; f->sigmoid = h.sig >= 0; f->cout = co;
      out->kind = FusedPlan::Tensor::kBlobF32;
      f->out_blob = out->blob;
      // fused-away values
      tensors[top_t[h.deconv][0]]->kind = FusedPlan::Tensor::kVirtual;
      tensors[top_t[h.skipconv][0]]->kind = FusedPlan::Tensor::kVirtual;
      tensors[top_t[h.crop][0]]->kind = FusedPlan::Tensor::kVirtual;
      if (h.sig >= 0) tensors[top_t[h.elt][0]]->kind = FusedPlan::Tensor::kVirtual;
      done[h.deconv] = done[h.skipconv] = done[h.crop] = done[h.elt] = 1;
      if (h.sig >= 0) done[h.sig] = 1;
      off += co;
    }
    return true;
  }

  // A convolution that will be absorbed by a head group (1x1 + bias whose output feeds a Crop reference)
  bool IsHeadSkipConv(int i) {
    ConvBase* c = As<ConvBase>(net.layers()[i].get());
    if (!c->bias_term()) return false;
    for (int cons : tensors[top_t[i][0]]->consumers)
      if (cons >= 0 && std::string(net.layers()[cons]->type()) == "Crop") return true;
    return false;
  }

  bool Run() {
    BuildDataflow();
    const int nl = static_cast<int>(net.layers().size());
    for (int i = 0; i < nl; ++i) {
      if (done[i]) continue;
      const std::string type = net.layers()[i]->type();
      if (type == "Split") { done[i] = 1; continue; }
      if (type == "Convolution") {
        if (IsHeadSkipConv(i)) continue;         // picked up by MatchHeads
        if (!MatchConv(i)) return false;
      } else if (type == "Eltwise") {
        if (!MatchEltwise(i)) return false;
      } else if (type == "Pooling") {
        if (!MatchPool(i)) return false;
      } else if (type == "Deconvolution") {
        if (!MatchHeads(i)) return false;
      } else {
        return Fail("layer " + net.layer_names()[i] + " (" + type + ") is not part of a fusable pattern");
      }
    }
    for (int i = 0; i < nl; ++i)
      if (!done[i]) return Fail("layer " + net.layer_names()[i] + " was left unmatched");
    return true;
  }
};

size_t AlignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

PlanWeightCache::~PlanWeightCache() {
  for (void* p : allocs) dc_free(p);
}
bool PlanWeightCache::Stale() const {
  for (const Seen& we : epochs_)
    if (we.blob->data() != we.mem || we.mem->host_write_epoch() != we.epoch) return true;
  return false;
}
void PlanWeightCache::Snapshot(Net<float>& net) {
  epochs_.clear();
  for (const auto& layer : net.layers())
    for (const auto& blob : layer->blobs()) {
      blob->cpu_data();     // make sure the SyncedMemory exists and is host-readable
      Seen seen = {blob.get(), blob->data(), blob->data()->host_write_epoch()};
      epochs_.push_back(seen);
    }
}

FusedPlan* FusedPlan::Build(Net<float>& net, bool materialize, std::string* why_not, std::shared_ptr<PlanWeightCache>* cache, bool dry_run) {
  FusedPlan* plan = new FusedPlan();
  if (!dry_run && cache && *cache && !(*cache)->Stale()) plan->weights_ = *cache;
  plan->net_ = &net;
  plan->materialize_ = materialize;
  std::string why;
  if (!plan->Match(net, materialize, &why)) {
    if (why_not) *why_not = why;
    delete plan;
    return nullptr;
  }
  if (why_not) why_not->clear();
  plan->PlanSchedule();
  plan->PlanMemory(dry_run);
  const std::string bad = plan->VerifySchedule();
  if (!bad.empty()) LOG(FATAL) << "fused plan: inconsistent schedule: " << bad;
  if (dry_run) return plan;
  plan->UploadWeights(net);
  if (cache) *cache = plan->weights_;
  return plan;
}

bool FusedPlan::Match(Net<float>& net, bool materialize, std::string* why) {
  Matcher m(net, tensors_, steps_);
  if (!m.Run()) { *why = m.why; return false; }
  // Net outputs that ended up as split activations, and (on request) every named intermediate,
  // are converted back into their fp32 NCHW blobs.
  std::set<int> out_blobs(net.output_blob_indices().begin(), net.output_blob_indices().end());
  std::vector<Step*> with_copies;
  for (Step* s : steps_) {
    with_copies.push_back(s);
    Tensor* t = s->out;
    if (t && t->kind == Tensor::kSplit && t->blob >= 0 && (materialize || out_blobs.count(t->blob))) {
      Step* c = new Step();
      c->type = Step::kToBlob;
      c->name = net.blob_names()[t->blob] + "/to_blob";
      c->in = t;
      c->out_blob = t->blob;
      with_copies.push_back(c);
    }
  }
  steps_.swap(with_copies);
  if (materialize)
    for (size_t i = 0; i < net.layers().size(); ++i)
      if (std::string(net.layers()[i]->type()) == "Split") split_layers_.push_back(static_cast<int>(i));
  return true;
}

namespace {
size_t EnvMiB(const char* name, size_t dflt) {
  const char* e = getenv(name);
  return (e && *e) ? static_cast<size_t>(atof(e) * 1048576.0) : dflt;
}
bool EnvOn(const char* name, bool dflt) {
  const char* e = getenv(name);
  return (e && *e) ? e[0] != '0' : dflt;
}
}  // namespace

// Step-order liveness, then the L2-resident schedule: the steps are grouped into SEGMENTS -- the bottleneck blocks of one
// ResNet stage, i.e. consecutive ConvBN steps whose block outputs share one geometry -- and a segment whose per-image working
// set (the block input/shortcut, which the block output overwrites in place, plus the two narrow intermediates) fits the L2
// budget several times over is executed sub-batch by sub-batch: all its steps for images [0, c), then for [c, 2c), ...  Within
// a pass every activation a conv reads was written a few launches earlier by the same pass and is still in the 126 MB L2, and
// the tensors that never leave the segment are allocated for c images only, so successive passes overwrite the same lines
// instead of streaming 16 images' worth of each tensor through HBM (75 GB per-conv vs 33 GB block-fused).
// DC_L2_CHUNK_MB sets the budget (0 = off); DC_CHUNK_PLAN="c0,c1,.." forces the images per pass of the candidate segments in order.
void FusedPlan::PlanSchedule() {
  const int ns = static_cast<int>(steps_.size());
  for (int s = 0; s < ns; ++s) {
    Step* st = steps_[s];
    if (st->ws) { st->ws->def_step = s; st->ws->last_step = s; }
    Tensor* ins[3] = {st->in, st->in2, st->col};
    for (Tensor* t : ins)
      if (t) t->last_step = std::max(t->last_step, s);
    if (st->out && st->out->def_step < 0) {
      st->out->def_step = s;
      st->out->last_step = std::max(st->out->last_step, s);
    }
  }
  for (Tensor* t : tensors_) { t->alloc_step = t->def_step; t->free_step = t->last_step; }

  // Off by default: measured on B200 (profiles/r2_chunk_sweep.md) the chunked schedule cuts a forward's DRAM traffic from 74.6 GB
  // to 32.8 GB, yet the step gets SLOWER -- operand delivery is already hidden behind the MMAs (profiles/r2_microbench_bound.txt), so
  // fewer DRAM bytes buy nothing, while a sub-batch launch has 3-16x fewer tiles per wave and pays ramp + tail as many times more often.
  const size_t budget = EnvMiB("DC_L2_CHUNK_MB", 0);
  std::vector<int> forced;
  if (const char* e = getenv("DC_CHUNK_PLAN")) {
    std::stringstream ss(e);
    std::string tok;
    while (std::getline(ss, tok, ',')) forced.push_back(atoi(tok.c_str()));
  }
  const bool inplace = EnvOn("DC_INPLACE_RESIDUAL", true);
  auto chunkable = [&](const Step* st) {
    return st->type == Step::kConvBN && st->in->kind == Tensor::kSplit && st->out->kind == Tensor::kSplit && st->in->n > 1 &&
           st->out->n == st->in->n && (!st->in2 || (st->in2->kind == Tensor::kSplit && st->in2->n == st->in->n));
  };
  size_t candidate = 0;
  auto close_segment = [&](int a, int b) {
    if (b <= a) return;
    const int N = steps_[a]->in->n;
    auto local = [&](const Tensor* t) { return t && t->bytes == 0 && t->kind == Tensor::kSplit && t->def_step >= a && t->last_step <= b; };
    // per-image working set: the largest sum of segment-local tensors alive at one step (a block output that takes over its
    // shortcut's storage counts once)
    size_t ws = 0;
    for (int j = a; j <= b; ++j) {
      size_t live = 0;
      for (const Tensor* t : tensors_)
        if (local(t) && t->def_step <= j && j <= t->last_step) live += t->elems() / N * 4;
      const Step* st = steps_[j];
      if (inplace && st->in2 && local(st->in2) && local(st->out) && st->in2->last_step == j && st->in2 != st->in && st->in2->elems() == st->out->elems())
        live -= st->out->elems() / N * 4;
      ws = std::max(ws, live);
    }
    int chunk = N;
    if (candidate < forced.size()) chunk = forced[candidate] > 0 ? std::min(forced[candidate], N) : N;
    else if (budget > 0 && ws > 0 && ws <= budget) chunk = static_cast<int>(std::min<size_t>(N, budget / ws));
    ++candidate;
    if (chunk < N) {                       // even passes: 16 images at 5 per pass -> 4 passes of 4
      const int passes = (N + chunk - 1) / chunk;
      chunk = (N + passes - 1) / passes;
    }
    Segment seg = {a, b, chunk, ws};
    segments_.push_back(seg);
    if (chunk >= N) return;
    // DC_PLAN_BREAK_LIVENESS=1 (tests only) leaves out the widening below, which VerifySchedule must then reject
    const bool widen = !EnvOn("DC_PLAN_BREAK_LIVENESS", false);
    for (Tensor* t : tensors_) {
      if (local(t)) { t->chunk_n = chunk; continue; }
      if (t->def_step < 0 || !widen) continue;
      // a tensor that crosses the segment boundary must own its storage for the WHOLE segment: pass k+1 still reads the
      // segment inputs after pass k has written the segment outputs
      if (t->def_step < a && t->last_step >= a && t->last_step <= b) t->free_step = std::max(t->free_step, b);
      if (t->def_step >= a && t->def_step <= b && t->last_step > b) t->alloc_step = std::min(t->alloc_step, a);
    }
  };
  for (int s = 0; s < ns;) {
    if (!chunkable(steps_[s])) { ++s; continue; }
    int e = s;
    while (e + 1 < ns && chunkable(steps_[e + 1]) && steps_[e + 1]->in->n == steps_[s]->in->n) ++e;
    // blocks end at the convs that absorb a shortcut; consecutive blocks with the same output geometry form a segment
    int seg_first = s, block_first = s;
    long long key = -1;
    for (int j = s; j <= e; ++j) {
      if (!steps_[j]->in2) continue;
      const Tensor* o = steps_[j]->out;
      const long long k = (static_cast<long long>(o->h) * 65536 + o->w) * 65536 + o->c;
      if (key >= 0 && k != key) { close_segment(seg_first, block_first - 1); seg_first = block_first; }
      key = k;
      block_first = j + 1;
    }
    close_segment(seg_first, e);
    s = e + 1;
  }
  // the launch order
  std::vector<int> seg_of(ns, -1);
  for (size_t g = 0; g < segments_.size(); ++g)
    for (int j = segments_[g].first; j <= segments_[g].last; ++j) seg_of[j] = static_cast<int>(g);
  for (int s = 0; s < ns;) {
    const int g = seg_of[s];
    if (g < 0 || segments_[g].chunk >= steps_[s]->in->n) {
      Issue is = {s, 0, 0};
      schedule_.push_back(is);
      ++s;
      continue;
    }
    const Segment& seg = segments_[g];
    const int N = steps_[s]->in->n;
    for (int i0 = 0; i0 < N; i0 += seg.chunk)
      for (int j = seg.first; j <= seg.last; ++j) {
        Issue is = {j, i0, std::min(seg.chunk, N - i0)};
        schedule_.push_back(is);
      }
    s = seg.last + 1;
  }
}

// Liveness-based placement of every arena tensor (first-fit over a sorted free list).
void FusedPlan::PlanMemory(bool dry_run) {
  struct Free { size_t off, size; };
  std::vector<Free> free_list;
  size_t top = 0;
  auto alloc = [&](size_t bytes) {
    bytes = AlignUp(bytes, 1024);
    for (size_t i = 0; i < free_list.size(); ++i) {
      if (free_list[i].size >= bytes) {
        const size_t off = free_list[i].off;
        free_list[i].off += bytes;
        free_list[i].size -= bytes;
        if (free_list[i].size == 0) free_list.erase(free_list.begin() + i);
        return off;
      }
    }
    // grow: extend a trailing free block if there is one
    if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
      const size_t off = free_list.back().off;
      top = off + bytes;
      free_list.pop_back();
      return off;
    }
    const size_t off = top;
    top += bytes;
    return off;
  };
  auto release = [&](size_t off, size_t bytes) {
    bytes = AlignUp(bytes, 1024);
    Free f = {off, bytes};
    auto it = std::lower_bound(free_list.begin(), free_list.end(), f, [](const Free& a, const Free& b) { return a.off < b.off; });
    it = free_list.insert(it, f);
    const size_t i = it - free_list.begin();
    if (i + 1 < free_list.size() && free_list[i].off + free_list[i].size == free_list[i + 1].off) {
      free_list[i].size += free_list[i + 1].size;
      free_list.erase(free_list.begin() + i + 1);
    }
    if (i > 0 && free_list[i - 1].off + free_list[i - 1].size == free_list[i].off) {
      free_list[i - 1].size += free_list[i].size;
      free_list.erase(free_list.begin() + i);
    }
  };
  auto size_of = [](const Tensor* t) -> size_t {
    if (t->kind == Tensor::kRaw) return t->raw_bytes;
    if (t->kind == Tensor::kF32Rows) return static_cast<size_t>(t->n) * t->h * t->w * t->ld * 4;
    if (t->kind != Tensor::kSplit) return 0;
    return t->chunk_n > 0 ? t->elems() / t->n * t->chunk_n * 4 : t->elems() * 4;
  };
  const bool inplace = EnvOn("DC_INPLACE_RESIDUAL", true);
  for (size_t s = 0; s < steps_.size(); ++s) {
    const int si = static_cast<int>(s);
    // In place: a block's output takes over the storage of the shortcut it absorbs when this step is the shortcut's last
    // reader (every element is read, then written, by the same epilogue warp; the conv's own input is another tensor).
    // Halves the block's resident footprint and turns the output's write misses into hits on lines the read just brought in.
    Step* st = steps_[s];
    Tensor* taken = nullptr;
    if (inplace && st->type == Step::kConvBN && st->in2 && st->out && st->in2 != st->in && st->in2->bytes && !st->in2->gave_memory &&
        st->in2->free_step == si && st->out->alloc_step == si && st->out->bytes == 0 && size_of(st->out) == st->in2->bytes &&
        st->in2->chunk_n == st->out->chunk_n) {
      taken = st->in2;
      st->out->bytes = taken->bytes;
      st->out->offset = taken->offset;
      taken->gave_memory = true;
    }
    for (Tensor* t : tensors_) {
      if (t->alloc_step != si || t->bytes) continue;
      const size_t b = size_of(t);
      if (!b) continue;
      t->bytes = b;
      t->offset = alloc(b);
    }
    for (Tensor* u : tensors_)
      if (u->bytes && u->free_step == si && !u->gave_memory) release(u->offset, u->bytes);
  }
  // the tail of the arena is the split-K scratch of this plan's under-filled conv launches (dc_conv_args.splitk_workspace):
  // one region for the whole plan, the launches that use it are serialised on the plan's stream
  splitk_ws_bytes_ = dc_splitk_workspace_bytes();
  const size_t ws_off = AlignUp(std::max<size_t>(top, 1024), 1024);
  arena_bytes_ = ws_off + splitk_ws_bytes_;
  if (dry_run) return;
  DC_CHECK(dc_malloc(&arena_, arena_bytes_));
  splitk_ws_ = static_cast<char*>(arena_) + ws_off;
  for (Tensor* t : tensors_)
    if (t->bytes) t->ptr = static_cast<char*>(arena_) + t->offset;
}

std::string FusedPlan::VerifySchedule() const {
  // ownership of arena bytes: interval start -> (end, tensor id, image or -1 for an unsliced tensor)
  struct Own { size_t end; int tensor, image; };
  std::map<size_t, Own> own;
  auto write = [&](size_t b, size_t e, int tensor, int image) {
    if (b >= e) return;
    auto it = own.lower_bound(b);
    if (it != own.begin()) {
      auto pr = std::prev(it);
      if (pr->second.end > b) {                       // split the interval that straddles b
        Own tail = pr->second;
        pr->second.end = b;
        if (tail.end > e) own[e] = tail;
      }
    }
    it = own.lower_bound(b);
    while (it != own.end() && it->first < e) {
      if (it->second.end > e) { Own tail = it->second; own.erase(it); own[e] = tail; break; }
      it = own.erase(it);
    }
    Own o = {e, tensor, image};
    own[b] = o;
  };
  auto owned = [&](size_t b, size_t e, int tensor, int image) {
    size_t pos = b;
    auto it = own.upper_bound(b);
    if (it != own.begin()) --it;
    for (; it != own.end() && pos < e; ++it) {
      if (it->second.end <= pos) continue;
      if (it->first > pos || it->second.tensor != tensor || it->second.image != image) return false;
      pos = it->second.end;
    }
    return pos >= e;
  };
  // visits the arena ranges of images [i0, i0 + cn) of tensor t (both planes of a split tensor)
  auto ranges = [&](const Tensor* t, int i0, int cn, const std::function<bool(size_t, size_t, int)>& fn) {
    if (!t || !t->bytes) return true;
    if (t->kind != Tensor::kSplit) return fn(t->offset, t->offset + t->bytes, -1);
    const size_t img = t->elems() / t->n * 2;
    const int first = cn > 0 ? i0 : 0, count = cn > 0 ? cn : t->n;
    const bool local = t->chunk_n > 0;
    if (local && (cn == 0 || cn > t->chunk_n)) return false;
    const size_t plane = local ? static_cast<size_t>(count) * img : static_cast<size_t>(t->n) * img;
    for (int k = 0; k < count; ++k) {
      const size_t slot = local ? k : first + k;
      if (!fn(t->offset + slot * img, t->offset + (slot + 1) * img, first + k)) return false;
      if (!fn(t->offset + plane + slot * img, t->offset + plane + (slot + 1) * img, first + k)) return false;
    }
    return true;
  };
  for (const Issue& is : schedule_) {
    const Step* st = steps_[is.step];
    const Tensor* reads[3] = {st->in, st->type == Step::kConvBN || st->type == Step::kHeadFinish ? st->in2 : nullptr, st->col};
    for (const Tensor* t : reads) {
      if (!t || !t->bytes) continue;
      if (t->bytes > arena_bytes_ || t->offset + t->bytes > arena_bytes_) return "tensor outside the arena at step " + st->name;
      if (!ranges(t, is.i0, is.cn, [&](size_t b, size_t e, int image) { return owned(b, e, t->id, image); }))
        return "step " + st->name + " (images from " + std::to_string(is.i0) + ") reads tensor " + std::to_string(t->id) + " after its storage was reused";
    }
    if (st->ws) write(st->ws->offset, st->ws->offset + st->ws->bytes, st->ws->id, -1);
    const Tensor* o = st->out;
    if (o && o->bytes) {
      if (o->offset + o->bytes > arena_bytes_) return "tensor outside the arena at step " + st->name;
      if (!ranges(o, is.i0, is.cn, [&](size_t b, size_t e, int image) { write(b, e, o->id, image); return true; }))
        return "step " + st->name + " writes a sub-batch larger than its segment-local tensor";
    }
  }
  return std::string();
}

bool FusedPlan::WeightsStale() const { return !weights_ || weights_->Stale(); }

void FusedPlan::UploadWeights(Net<float>& net) {
  void* stream = Caffe::stream();
  const bool reuse = static_cast<bool>(weights_);
  if (!reuse) {
    weights_.reset(new PlanWeightCache());
    weights_->Snapshot(net);
  }
  PlanWeightCache& wc = *weights_;
  // steps are keyed by type + name: the same layers fuse the same way whatever the input shape
  auto key_of = [](const Step* st) { return std::to_string(static_cast<int>(st->type)) + (st->stem_tc ? "t:" : ":") + st->name; };
  // A cache that survived (same weights, new input shape or new option such as the skipped-outputs set) serves every step it has an
  // entry for; steps it has never seen (another head grouping) are packed below and added to it.
  auto upload = [&](const void* host, size_t bytes) {
    void* d = nullptr;
    DC_CHECK(dc_malloc(&d, bytes));
    wc.allocs.push_back(d);
    wc.bytes += bytes;
    DC_CHECK(dc_memcpy_async(d, host, bytes, DC_H2D, stream));
    DC_CHECK(dc_stream_sync(stream));      // the host staging buffer is reused right after
    return d;
  };
  // per-channel affine of an optional BatchNorm + Scale pair
  auto fold = [&](int bn_layer, int scale_layer, int channels, std::vector<float>* a, std::vector<float>* b) {
    a->assign(channels, 1.f);
    b->assign(channels, 0.f);
    const float* gamma = nullptr;
    const float* beta = nullptr;
    if (scale_layer >= 0) {
      Layer<float>* sl = net.layers()[scale_layer].get();
      gamma = sl->blobs()[0]->cpu_data();
      if (sl->blobs().size() > 1) beta = sl->blobs()[1]->cpu_data();
    }
    if (bn_layer >= 0) {
      BatchNormLayer<float>* bl = As<BatchNormLayer<float> >(net.layers()[bn_layer].get());
      DC_CHECK(dc_fold_bn_scale(bl->blobs()[0]->cpu_data(), bl->blobs()[1]->cpu_data(), bl->blobs()[2]->cpu_data()[0], bl->eps(), gamma, beta,
                                channels, a->data(), b->data()));
    } else if (gamma) {
      for (int c = 0; c < channels; ++c) { (*a)[c] = gamma[c]; (*b)[c] = beta ? beta[c] : 0.f; }
    }
  };
  for (Step* st : steps_) {
    if (reuse) {
      auto it = wc.entries.find(key_of(st));
      if (it != wc.entries.end()) {
        st->w_dev = it->second.w; st->scale_dev = it->second.scale; st->shift_dev = it->second.shift;
        continue;
      }
    }
    if (st->type == Step::kConv1) {
      Layer<float>* cl = net.layers()[st->conv_layer].get();
      std::vector<float> a, b;
      fold(st->bn_layer, st->scale_layer, 64, &a, &b);
      if (st->stem_tc) {
        std::vector<uint16_t> packed(2 * 64 * 256);
        std::vector<float> rs(64);
        DC_CHECK(dc_pack_conv1_tc_weight(cl->blobs()[0]->cpu_data(), packed.data(), rs.data()));
        for (int c = 0; c < 64; ++c) a[c] *= rs[c];
        st->w_dev = upload(packed.data(), packed.size() * 2);
      } else {
        std::vector<float> wp(147 * 64);
        DC_CHECK(dc_pack_conv1_weight(cl->blobs()[0]->cpu_data(), wp.data()));
        st->w_dev = upload(wp.data(), wp.size() * 4);
      }
      st->scale_dev = static_cast<float*>(upload(a.data(), 64 * 4));
      st->shift_dev = static_cast<float*>(upload(b.data(), 64 * 4));
    } else if (st->type == Step::kConvBN) {
      ConvBase* cl = As<ConvBase>(net.layers()[st->conv_layer].get());
      const int cout = st->cout, cin = cl->channels(), rows = dc_packed_rows(cout);
      const size_t K = static_cast<size_t>(st->kh) * st->kw * cin;
      std::vector<uint16_t> packed(2 * rows * K);
      std::vector<float> rs(rows), a, b, scale(rows, 1.f), shift(rows, 0.f);
      DC_CHECK(dc_pack_conv_weight(cl->blobs()[0]->cpu_data(), cout, cin, st->kh, st->kw, packed.data(), rs.data()));
      fold(st->bn_layer, st->scale_layer, cout, &a, &b);
      for (int c = 0; c < cout; ++c) {
        scale[c] = a[c] * rs[c];
        shift[c] = b[c] + (cl->bias_term() ? cl->blobs()[1]->cpu_data()[c] : 0.f);
      }
      st->w_dev = upload(packed.data(), packed.size() * 2);
      st->scale_dev = static_cast<float*>(upload(scale.data(), rows * 4));
      st->shift_dev = static_cast<float*>(upload(shift.data(), rows * 4));
    } else if (st->type == Step::kHeadGemm) {
      // concatenate the heads' weight blobs along the output-channel axis, then pack once
      const int cin = st->in->c;
      int ctot = 0;
      for (int l : st->merged_layers) ctot += As<ConvBase>(net.layers()[l].get())->num_output();
      const int taps = st->deconv_rows ? 9 : 1;
      const int real = ctot;
      if (!st->deconv_rows) ctot = std::max(ctot, st->cout);        // skip matrix zero-padded to a 128-row tile (MatchHeads)
      std::vector<float> wcat(static_cast<size_t>(cin) * ctot * taps, 0.f);
      std::vector<float> bias(ctot, 0.f);
      int off = 0;
      for (int l : st->merged_layers) {
        ConvBase* cl = As<ConvBase>(net.layers()[l].get());
        const int co = cl->num_output();
        const float* w = cl->blobs()[0]->cpu_data();
        if (st->deconv_rows) {       // W[ci][co][3][3] -> Wcat[ci][off+co][3][3]
          for (int ci = 0; ci < cin; ++ci)
            memcpy(&wcat[(static_cast<size_t>(ci) * ctot + off) * 9], w + static_cast<size_t>(ci) * co * 9, sizeof(float) * co * 9);
        } else {                      // W[co][ci] -> Wcat[off+co][ci]
          memcpy(&wcat[static_cast<size_t>(off) * cin], w, sizeof(float) * co * cin);
        }
        if (cl->bias_term())
          for (int c = 0; c < co; ++c) bias[off + c] = cl->blobs()[1]->cpu_data()[c];
        off += co;
      }
      const int rows = dc_packed_rows(ctot * taps);
      std::vector<uint16_t> packed(2 * static_cast<size_t>(rows) * cin);
      std::vector<float> rs(rows), shift(rows, 0.f);
      if (st->deconv_rows) DC_CHECK(dc_pack_deconv_weight(wcat.data(), cin, ctot, 3, 3, packed.data(), rs.data()));
      else DC_CHECK(dc_pack_conv_weight(wcat.data(), ctot, cin, 1, 1, packed.data(), rs.data()));
      st->w_dev = upload(packed.data(), packed.size() * 2);
      st->scale_dev = static_cast<float*>(upload(rs.data(), rows * 4));
      if (!st->deconv_rows) {
        // both biases of a head land on the same output element: fold the deconvolution's into the
        // skip GEMM's shift (the deconv GEMM step precedes this one and recorded its biases there)
        for (int c = 0; c < real; ++c) shift[c] = bias[c];
        for (Step* other : steps_)
          if (other->type == Step::kHeadGemm && other->deconv_rows) {
            int o2 = 0;
            for (int l : other->merged_layers) {
              ConvBase* dl = As<ConvBase>(net.layers()[l].get());
              if (dl->bias_term())
                for (int c = 0; c < dl->num_output(); ++c) shift[o2 + c] += dl->blobs()[1]->cpu_data()[c];
              o2 += dl->num_output();
            }
          }
      }
      st->shift_dev = static_cast<float*>(upload(shift.data(), rows * 4));
    }
    if (st->w_dev) {
      PlanWeightCache::Entry e;
      e.w = st->w_dev; e.scale = st->scale_dev; e.shift = st->shift_dev;
      wc.entries[key_of(st)] = e;
    }
  }
}

// Blob device pointers the steps read/write, resolved OUTSIDE any graph capture (gpu_data() may upload the
// input; overwrite_gpu_data() may allocate): one entry per step, nullptr when the step touches no blob.
void FusedPlan::Run() {
  Net<float>& net = *net_;
  void* stream = Caffe::stream();
  std::vector<const void*> ptrs(steps_.size(), nullptr);
  for (size_t i = 0; i < steps_.size(); ++i) {
    const Step* st = steps_[i];
    if (st->type == Step::kConv1) ptrs[i] = net.blobs()[st->in->blob]->gpu_data();
    else if (st->type == Step::kHeadFinish || st->type == Step::kToBlob) ptrs[i] = net.blobs()[st->out_blob]->overwrite_gpu_data();
  }
  static const bool graphs_on = [] { const char* e = getenv("DC_CUDA_GRAPH"); return !(e && e[0] == '0'); }();
  if (graphs_on && !step_timing_ && !graph_failed_) {
    if (graph_ != nullptr && ptrs != graph_ptrs_) { dc_graph_destroy(graph_); graph_ = nullptr; }
    if (graph_ == nullptr) {
      if (dc_graph_begin(stream) == 0) {
        try {
          IssueSteps(ptrs, stream);
        } catch (...) {
          // a launch was refused mid-capture (FatalError under the Python binding): leave capture mode before the error
          // travels on, or every later call on this stream would fail with "operation not permitted when stream is capturing"
          void* dead = nullptr;
          if (dc_graph_end(stream, &dead) == 0 && dead) dc_graph_destroy(dead);
          graph_failed_ = true;
          throw;
        }
        if (dc_graph_end(stream, &graph_) != 0) { graph_ = nullptr; graph_failed_ = true; LOG(WARNING) << "CUDA graph capture failed (" << DcLastError() << "); launching step by step"; }
        else graph_ptrs_ = ptrs;
      } else {
        graph_failed_ = true;
      }
    }
    if (graph_ != nullptr) {
      DC_CHECK(dc_graph_launch(graph_, stream));
      for (int l : split_layers_) net.layers()[l]->Forward(net.bottom_vecs()[l], net.top_vecs()[l]);
      return;
    }
  }
  IssueSteps(ptrs, stream);
  for (int l : split_layers_) net.layers()[l]->Forward(net.bottom_vecs()[l], net.top_vecs()[l]);
}

void FusedPlan::IssueSteps(const std::vector<const void*>& blob_ptrs, void* stream) {
  Net<float>& net = *net_;
  if (step_timing_ && events_.size() != schedule_.size() + 1) {
    for (void* e : events_) dc_event_destroy(e);
    events_.assign(schedule_.size() + 1, nullptr);
    for (void*& e : events_) DC_CHECK(dc_event_create(&e));
  }
  // address of images [i0, ..) of a split tensor + the hi->lo plane distance a sub-batch launch must be told (0 = dense)
  auto slice = [](const Tensor* t, int i0, int cn, long long* plane) -> void* {
    *plane = 0;
    if (cn == 0 || t->chunk_n > 0) return t->ptr;          // whole batch, or a segment-local tensor (holds this pass only)
    *plane = static_cast<long long>(t->elems());
    return static_cast<char*>(t->ptr) + static_cast<size_t>(i0) * (t->elems() / t->n) * 2;
  };
  size_t issue_index = 0;
  // Weight tiles stay in L2 with the evict_last priority (DC_WEIGHTS_EVICT_LAST=0 disables): the one L2 hint of round 2's sweep that
  // measured a gain (res4's 3x3 convs 7.5 -> 7.1 ms per 16x720p step).  evict_first on streamed activations, evict_last on small
  // outputs, a serpentine tile order and merged accumulators were measured too and removed again: profiles/r2_chunk_sweep.md,
  // tools/experiments/r2_removed_switches.patch.
  const bool weights_last = EnvOn("DC_WEIGHTS_EVICT_LAST", true);
  for (const Issue& is : schedule_) {
    Step* st = steps_[is.step];
    if (step_timing_) DC_CHECK(dc_event_record(events_[issue_index], stream));
    const void* bp = blob_ptrs[is.step];
    ++issue_index;
    switch (st->type) {
      case Step::kConv1: {
        const float* x = static_cast<const float*>(bp);
        if (st->stem_tc)
          DC_CHECK(dc_conv1_tc_forward(x, st->in->n, st->in->h, st->in->w, st->w_dev, st->scale_dev, st->shift_dev, st->ws->ptr, st->out->ptr, stream));
        else
          DC_CHECK(dc_conv1_forward(x, st->in->n, st->in->h, st->in->w, static_cast<const float*>(st->w_dev), st->scale_dev, st->shift_dev,
                                    st->out->ptr, stream));
        break;
      }
      case Step::kConvBN:
      case Step::kHeadGemm: {
        dc_conv_args a;
        memset(&a, 0, sizeof(a));
        const bool conv = st->type == Step::kConvBN;
        a.x = slice(st->in, is.i0, is.cn, &a.x_plane);
        a.n = is.cn > 0 ? is.cn : st->in->n; a.h = st->in->h; a.w = st->in->w; a.cin = st->in->c;
        a.cout = st->cout; a.kh = st->kh; a.kw = st->kw; a.pad = st->pad; a.dilation = st->dil;
        a.w_packed = st->w_dev; a.scale = st->scale_dev; a.shift = st->shift_dev;
        a.residual = st->in2 && conv ? slice(st->in2, is.i0, is.cn, &a.residual_plane) : nullptr;
        a.relu = st->relu;
        a.out_f32_rows = conv ? 0 : 2;
        a.ldc = st->out->ld;
        a.out = conv ? slice(st->out, is.i0, is.cn, &a.out_plane) : st->out->ptr;
        a.stride = conv ? st->stride : 1;
        a.splitk_workspace = splitk_ws_;
        a.splitk_workspace_bytes = splitk_ws_bytes_;
        a.weights_evict_last = weights_last ? 1 : 0;
        DC_CHECK(dc_conv_forward(&a, stream));
        break;
      }
      case Step::kSubsample:
        DC_CHECK(dc_subsample_forward(st->in->ptr, st->in->n, st->in->h, st->in->w, st->in->c, st->stride, st->out->ptr, stream));
        break;
      case Step::kMaxPool:
        DC_CHECK(dc_maxpool_forward(st->in->ptr, st->in->n, st->in->h, st->in->w, st->in->c, st->pool_k, st->pool_s, st->out->ptr, stream));
        break;
      case Step::kHeadFinish: {
        Blob<float>* ob = net.blobs()[st->out_blob].get();
        DC_CHECK(dc_head_finish(static_cast<const float*>(st->col->ptr), st->col->ld, st->col_off, static_cast<const float*>(st->in2->ptr),
                                st->in2->ld, st->skip_off, static_cast<float*>(const_cast<void*>(bp)), st->in->n, st->cout, st->in->h, st->in->w,
                                ob->height(), ob->width(), st->sigmoid, stream));
        break;
      }
      case Step::kToBlob:
        DC_CHECK(dc_split_to_nchw(st->in->ptr, st->in->n, st->in->c, st->in->h, st->in->w, static_cast<float*>(const_cast<void*>(bp)), stream));
        break;
    }
  }
  if (step_timing_) DC_CHECK(dc_event_record(events_[issue_index], stream));
}

std::vector<FusedPlan::StepInfo> FusedPlan::LastStepInfo() {
  static const char* kNames[] = {"Conv1", "ConvBN", "Subsample", "MaxPool", "HeadGemm", "HeadFinish", "ToBlob"};
  std::vector<StepInfo> out;
  // a step of a chunked segment is issued once per pass: its time is the sum over its launches
  std::vector<double> step_ms(steps_.size(), 0.0);
  if (events_.size() == schedule_.size() + 1)
    for (size_t k = 0; k < schedule_.size(); ++k) {
      float ms = 0.f;
      DC_CHECK(dc_event_elapsed_ms(events_[k], events_[k + 1], &ms));
      step_ms[schedule_[k].step] += ms;
    }
  for (size_t i = 0; i < steps_.size(); ++i) {
    const Step* st = steps_[i];
    StepInfo si;
    si.name = st->name;
    si.type = kNames[st->type];
    si.ms = step_ms[i];
    si.flops = 0;
    si.bytes = 0;
    auto tbytes = [](const Tensor* t) -> double {
      if (!t) return 0;
      if (t->kind == Tensor::kF32Rows) return 4.0 * t->n * t->h * t->w * t->ld;
      return 4.0 * t->elems();         // split fp16 hi+lo and fp32 blobs are both 4 B / element
    };
    switch (st->type) {
      case Step::kConv1: {
        si.flops = 2.0 * st->out->elems() * 147;
        si.bytes = tbytes(st->in) + tbytes(st->out) + 147 * 64 * 4;
        break;
      }
      case Step::kConvBN: {
        const double K = static_cast<double>(st->kh) * st->kw * st->in->c;
        si.flops = 2.0 * st->out->elems() * K;
        si.bytes = tbytes(st->in) / (static_cast<double>(st->stride) * st->stride) + tbytes(st->out) + tbytes(st->in2) + 4.0 * K * st->cout;
        break;
      }
      case Step::kHeadGemm: {
        const double pix = static_cast<double>(st->in->n) * st->in->h * st->in->w;
        si.flops = 2.0 * pix * st->cout * st->in->c;
        si.bytes = tbytes(st->in) + 4.0 * pix * st->cout + 4.0 * st->in->c * st->cout;
        break;
      }
      case Step::kHeadFinish: {
        Blob<float>* ob = net_->blobs()[st->out_blob].get();
        si.bytes = 4.0 * ob->count() * 2 + 4.0 * st->in->n * st->in->h * st->in->w * st->cout * 9;
        break;
      }
      case Step::kSubsample: si.bytes = 2 * tbytes(st->out); break;
      case Step::kMaxPool: si.bytes = tbytes(st->in) + tbytes(st->out); break;
      case Step::kToBlob: si.bytes = 2 * tbytes(st->in); break;
    }
    out.push_back(si);
  }
  return out;
}

std::vector<char> FusedPlan::WrittenBlobs() const {
  std::vector<char> w(net_->blobs().size(), 0);
  for (const Step* st : steps_)
    if (st->out_blob >= 0) w[st->out_blob] = 1;
  for (int l : split_layers_)             // Split tops alias their bottom (zero-copy), in layer order
    if (w[net_->bottom_ids(l)[0]])
      for (int t : net_->top_ids(l)) w[t] = 1;
  return w;
}

std::string FusedPlan::Describe() const {
  std::ostringstream s;
  static const char* kNames[] = {"Conv1", "ConvBN", "Subsample", "MaxPool", "HeadGemm", "HeadFinish", "ToBlob"};
  s << steps_.size() << " steps, " << schedule_.size() << " launch groups, arena " << (arena_bytes_ >> 20) << " MiB, weights " << (weight_bytes() >> 20) << " MiB\n";
  for (const Segment& g : segments_)
    s << "  segment " << steps_[g.first]->name << " .. " << steps_[g.last]->name << ": " << (g.bytes_per_image >> 10) << " KiB resident per image, "
      << g.chunk << " of " << steps_[g.first]->in->n << " images per pass\n";
  if (EnvOn("DC_DESCRIBE_SCHEDULE", false))      // one line per launch group, in launch order (profiling scripts map ncu rows to steps)
    for (const Issue& is : schedule_)
      s << "  issue " << kNames[steps_[is.step]->type] << " " << steps_[is.step]->name << " " << is.i0 << " " << is.cn << "\n";
  int inplace = 0;
  for (const Tensor* t : tensors_) inplace += t->gave_memory;
  s << "  " << inplace << " block outputs written in place over their shortcut\n";
  for (const Step* st : steps_) {
    s << "  " << kNames[st->type] << " " << st->name;
    if (st->in) s << " in=" << st->in->n << "x" << st->in->c << "x" << st->in->h << "x" << st->in->w;
    if (st->type == Step::kConvBN) s << " k" << st->kh << " p" << st->pad << " d" << st->dil << " ->" << st->cout << (st->relu ? " relu" : "") << (st->in2 ? " +res" : "");
    if (st->out && st->out->bytes) s << " out@" << st->out->offset << "+" << st->out->bytes << (st->out->chunk_n ? " (segment-local)" : "");
    s << "\n";
  }
  return s.str();
}

}  // namespace caffe
