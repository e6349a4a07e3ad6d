// ns_emu.cpp -- host SIMT emulation of the stream kernel (TEST PLUMBING, never shipped).
// Compiles crispy_b200/csrc/ns_kernel.cuh with NS_HOST_EMU: one OS thread per CUDA thread,
// pthread barriers for bar.sync / __syncthreads, a per-warp exchange buffer for shuffles.
// Lets `pytest -m "not gpu"` run the kernel's exact control flow against the oracle.
#define NS_HOST_EMU 1
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../crispy_b200/csrc/ns_host.h"
#include "../../crispy_b200/csrc/ns_kernel.cuh"

namespace ns {
thread_local EmuThread g_emu;
}

namespace {
template <int S>
struct Launch {
  ns::Params p;
  ns::CtaSmem<S> *sm;
  ns::EmuCta *cta;
};
template <int S>
struct ThreadArg {
  Launch<S> *l;
  int tid;
};
template <int S>
void *thread_main(void *a) {
  ThreadArg<S> *ta = (ThreadArg<S> *)a;
  ns::g_emu.cta = ta->l->cta;
  ns::g_emu.tid = ta->tid;
  ns::stream_kernel_body<S>(ta->l->p, *ta->l->sm);
  return nullptr;
}

template <int S>
int run(const ns::Params &p) {
  const int n_threads = S * ns::kGroupThreads;
  const int n_ctas = (p.n_streams + S - 1) / S;
  for (int c = 0; c < n_ctas; c++) {
    ns::EmuCta cta;
    std::vector<ns::EmuWarp> warps(n_threads / 32);
    cta.warps = warps.data();
    cta.cta_index = c;
    pthread_barrier_init(&cta.cta_bar, nullptr, n_threads);
    for (int i = 0; i < 16; i++) pthread_barrier_init(&cta.group_bar[i], nullptr, ns::kGroupThreads);
    for (auto &w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
    ns::CtaSmem<S> *sm = new ns::CtaSmem<S>();
    memset((void *)sm, 0xCD, sizeof(*sm));  // shared memory is not zeroed on a GPU either
    Launch<S> l{p, sm, &cta};
    std::vector<pthread_t> th(n_threads);
    std::vector<ThreadArg<S>> args(n_threads);
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 256 * 1024);
    for (int t = 0; t < n_threads; t++) {
      args[t] = ThreadArg<S>{&l, t};
      if (pthread_create(&th[t], &attr, thread_main<S>, &args[t]) != 0) return -2;
    }
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], nullptr);
    pthread_attr_destroy(&attr);
    delete sm;
    pthread_barrier_destroy(&cta.cta_bar);
    for (int i = 0; i < 16; i++) pthread_barrier_destroy(&cta.group_bar[i]);
    for (auto &w : warps) pthread_barrier_destroy(&w.bar);
  }
  return 0;
}
}  // namespace

extern "C" {
int ns_emu_state_floats(void) { return ns::kStateFloats; }
int ns_emu_dbg_floats(void) { return ns::kDbgFloats; }

// in/out/app use the same strides the device path uses; state is in/out ([n_streams][kStateFloats]).
int ns_emu_process(const void *model_blob, size_t model_len, const void *in, void *out, float *vad,
                   const float *app, float *state, float *dbg, int n_streams, int n_frames,
                   long long in_stride, long long out_stride, long long app_stride, int streams_per_cta,
                   unsigned flags, float volume, int out_frame_offset) {
  ns::Model m;
  std::string err;
  if (!ns::model_from_bytes(m, model_blob, model_len, err)) {
    fprintf(stderr, "ns_emu: %s\n", err.c_str());
    return -1;
  }
  ns::PackedRnn pk;
  ns::pack_rnn(m, pk);
  static ns::Tables tab;
  ns::make_tables(tab);
  ns::Params p;
  memset(&p, 0, sizeof(p));
  p.in = in;
  p.out = out;
  p.vad = vad;
  p.app = app;
  p.state = state;
  p.dbg = dbg;
  p.tables = &tab;
  p.rnn_hdr = &pk.hdr;
  p.rnn_words = pk.words.data();
  p.rnn_bias = pk.bias.data();
  p.in_stride = in_stride;
  p.out_stride = out_stride;
  p.vad_stride = n_frames;
  p.app_stride = app_stride;
  p.n_streams = n_streams;
  p.n_frames = n_frames;
  p.out_frame_offset = out_frame_offset;
  p.flags = flags;
  p.volume = volume;
  switch (streams_per_cta) {
    case 1: return run<1>(p);
    case 2: return run<2>(p);
    case 4: return run<4>(p);
    default: return -3;
  }
}
}
