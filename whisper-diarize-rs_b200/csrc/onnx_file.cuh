// onnx_file.cuh — dependency-free reader for the two ONNX files the crate loads through pyannote-rs:
// `segmentation-3.0.onnx` (pyannote_rs::get_segments, reference src/engine.rs:90, 117-122) and the WeSpeaker speaker-embedding
// export (EmbeddingExtractor::new, src/engine.rs:91, src/transcribe.rs:343).  ONNX Runtime is not linked: the file is walked as
// protobuf wire format (varint / length-delimited fields of ModelProto -> GraphProto -> NodeProto / TensorProto), which gives
//   * every tensor of the graph — initializers and Constant-node values — as fp32 (FLOAT, FLOAT16, DOUBLE; raw_data or typed fields),
//   * the node list in file (= topological) order with op types, input / output names and scalar / list attributes.
// The networks themselves are fixed architectures run by this library's kernels, so a loader only has to find each operator's
// parameters: it walks the nodes in order (InstanceNormalization, Conv, LSTM, MatMul + Add / Gemm, BatchNormalization) and takes
// the k-th operator of a kind for the k-th layer of that kind, checking every shape — exported names are often anonymous
// ("onnx::LSTM_789"), the operator sequence is not.  Host-only code (no CUDA calls).
#pragma once
#include <stdint.h>
#include <map>
#include <string>
#include <vector>

namespace wdr {

struct OnnxTensor {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = 0;            // TensorProto.DataType: 1 FLOAT, 6 INT32, 7 INT64, 10 FLOAT16, 11 DOUBLE
    std::vector<float> f32;   // values of a floating-point tensor
    std::vector<int64_t> i64; // values of an integer tensor (shape constants)
    int64_t count() const { int64_t n = 1; for (auto d : dims) n *= d; return n; }
};

struct OnnxAttr {
    std::string name;
    int64_t i = 0;
    float f = 0.0f;
    std::string s;
    std::vector<int64_t> ints;
    std::vector<float> floats;
    bool has_tensor = false;
    OnnxTensor t;
};

struct OnnxNode {
    std::string op_type, name;
    std::vector<std::string> inputs, outputs;
    std::vector<OnnxAttr> attrs;
    const OnnxAttr* attr(const char* n) const;
    int64_t attr_i(const char* n, int64_t dflt) const;
    float attr_f(const char* n, float dflt) const;
};

struct OnnxFile {
    std::string path;
    int64_t ir_version = 0;
    std::string producer;
    std::vector<OnnxNode> nodes;                 // file order (ONNX requires topological order)
    std::map<std::string, OnnxTensor> tensors;   // initializers + Constant-node outputs, by value name
    std::vector<std::string> inputs, outputs;    // graph inputs that are not initializers / graph outputs
    bool load(const char* path, std::string* err);
    bool parse(const unsigned char* data, size_t size, std::string* err);
    const OnnxTensor* tensor(const std::string& value_name) const;
};

// The parameters of the networks this library runs, gathered from an OnnxFile under the PyTorch-style names the seeded-weights path
// uses (so that one upload routine serves both): name -> fp32 values in PyTorch layout.
typedef std::map<std::string, std::vector<float>> NamedTensors;
// PyanNet (segmentation-3.0): wav_norm.{weight,bias}, conv{0,1,2}.weight, conv{1,2}.bias, norm{0,1,2}.{weight,bias},
// lstm.{weight_ih,weight_hh,bias_ih,bias_hh}_l{0..3}[_reverse] (PyTorch gate order i, f, g, o), linear{0,1}.{weight,bias},
// classifier.{weight,bias}.
bool onnx_extract_pyannet(const OnnxFile& f, NamedTensors* out, std::string* err);
// WeSpeaker ResNet34: <conv>.weight [co][ci][k][k] + <conv>.bias with the BatchNorm already folded (either by the exporter or here
// from a following BatchNormalization node), conv names conv1, layer<l>.<b>.conv1 / conv2 / shortcut in execution order (a block's shortcut after its conv2); seg_1.{weight,bias}.
// *emb_dim receives the embedding width (rows of the final Gemm).
bool onnx_extract_resnet34(const OnnxFile& f, NamedTensors* out, int* emb_dim, std::string* err);

}  // namespace wdr
