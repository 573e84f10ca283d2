// ggml_file.cuh — reader for whisper.cpp's `ggml-<model>.bin` checkpoints: the file
// `WhisperContext::new_with_params(model_path, ..)` opens (reference src/transcribe.rs:154; names at src/model_manager.rs:162).
//
// Layout (whisper.cpp `whisper_model_load`, legacy "ggml" container):
//   u32 magic 0x67676d6c | i32 n_vocab, n_audio_ctx, n_audio_state, n_audio_head, n_audio_layer, n_text_ctx, n_text_state,
//   n_text_head, n_text_layer, n_mels, ftype | i32 n_mel, n_fft, f32 filters[n_mel * n_fft] | i32 n_tokens, {u32 len, bytes}* |
//   tensors until EOF: i32 n_dims, i32 name_len, i32 type, i32 ne[n_dims] (innermost first), name, data.
// Tensor types 0 (f32) and 1 (f16) are read; quantised types are refused.  Host-only code (no CUDA calls).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>

namespace wdr {

struct GgmlTensorInfo {
    int type = 0;            // 0 f32, 1 f16
    std::vector<int> ne;     // ggml order: ne[0] is the fastest axis
    int64_t n_elem = 0;
    int64_t file_offset = 0;
};

struct SileroHeader {  // header of ggml-silero-v5.1.2.bin (src/model_manager.rs:305-315 downloads it; src/vad.rs:15-17 opens it)
    std::string model_type;
    int32_t version[3] = {0, 0, 0};
    int32_t n_encoder_layers = 0;
    std::vector<int32_t> enc_in, enc_out, enc_kernel;
    int32_t lstm_input = 0, lstm_hidden = 0, final_in = 0, final_out = 0;
};

struct GgmlFile {
    int32_t n_vocab = 0, n_audio_ctx = 0, n_audio_state = 0, n_audio_head = 0, n_audio_layer = 0;
    int32_t n_text_ctx = 0, n_text_state = 0, n_text_head = 0, n_text_layer = 0, n_mels = 0, ftype = 0;
    int32_t filt_n_mel = 0, filt_n_fft = 0;
    std::vector<float> filters;
    std::vector<std::string> tokens;
    std::map<std::string, GgmlTensorInfo> tensors;
    std::string path;
    // Parses the header, vocabulary and tensor index (tensor data stays on disk).  false + err on failure.
    bool open(const char* path, std::string* err);
    // The Silero VAD container: its own header, then the same tensor records.
    bool open_silero(const char* path, SileroHeader* h, std::string* err);
    bool read_tensor_index(FILE* f, std::string* err);
    // Reads one tensor as fp32 (f16 is widened).  false + err on failure (missing tensor, wrong element count, short read).
    bool read_f32(const std::string& name, int64_t expect_elems, std::vector<float>* out, std::string* err) const;
};

}  // namespace wdr
