// ns_host.h -- host-side (no CUDA) helpers: constant tables, RNN model blobs and their repacking
// into the kernel's job layout.  Shared by libcrispy_ns.so and by the host SIMT emulation.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "ns_common.h"

namespace ns {

void make_tables(Tables &t);

// The six layers of the RNNoise network, int8 as in nnnoiseless RnnModel.
struct DenseLayer {
  int nb_inputs = 0, nb_neurons = 0, activation = 0;
  std::vector<int8_t> weights, bias;  // weights[j*N + i]
};
struct GruLayer {
  int nb_inputs = 0, nb_neurons = 0, activation = 0;
  std::vector<int8_t> input_weights, recurrent_weights, bias;  // [j*3N + gate*N + i]
};
struct Model {
  DenseLayer input_dense;
  GruLayer vad_gru;
  DenseLayer vad_output;
  GruLayer noise_gru;
  GruLayer denoise_gru;
  DenseLayer denoise_output;
};

void model_synthetic(Model &m, uint64_t seed);
// "CRNSMDL1" binary blob or "rnnoise-nu model file version 1" text.  Returns false + err on failure.
bool model_from_bytes(Model &m, const void *blob, size_t len, std::string &err);
std::vector<uint8_t> model_to_bytes(const Model &m);

struct PackedRnn {
  RnnHeader hdr;
  std::vector<uint32_t> words;
  std::vector<float> bias;
};
void pack_rnn(const Model &m, PackedRnn &out);
// the same network for the tcgen05 recurrent core (ns_rnn_tc5.cuh): weight blocks as UMMA descriptors address them + biases
void pack_rnn_tc5(const Model &m, std::vector<uint8_t> &w, std::vector<float> &bias);

}  // namespace ns
