// ggml_file.cu — see ggml_file.cuh.  Host-only.
#include <string.h>
#include "ggml_file.cuh"

namespace wdr {

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

static float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {  // subnormal: normalise
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7F800000u | (man << 13);
    else bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float out;
    memcpy(&out, &bits, 4);
    return out;
}

bool GgmlFile::open(const char* p, std::string* err) {
    path = p ? p : "";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open " + path; return false; }
    auto fail = [&](const std::string& m) { *err = path + ": " + m; fclose(f); return false; };
    uint32_t magic = 0;
    if (!rd(f, &magic, 4) || magic != 0x67676d6cu) return fail("not a ggml whisper model (bad magic)");
    int32_t hp[11];
    if (!rd(f, hp, sizeof(hp))) return fail("truncated header");
    n_vocab = hp[0]; n_audio_ctx = hp[1]; n_audio_state = hp[2]; n_audio_head = hp[3]; n_audio_layer = hp[4];
    n_text_ctx = hp[5]; n_text_state = hp[6]; n_text_head = hp[7]; n_text_layer = hp[8]; n_mels = hp[9]; ftype = hp[10];
    if (n_vocab <= 0 || n_vocab > 200000 || n_audio_state <= 0 || n_audio_state > 8192 || n_audio_layer <= 0 || n_audio_layer > 128 ||
        n_text_layer <= 0 || n_text_layer > 128 || n_mels <= 0 || n_mels > 512)
        return fail("implausible hyper-parameters");
    if (!rd(f, &filt_n_mel, 4) || !rd(f, &filt_n_fft, 4) || filt_n_mel <= 0 || filt_n_fft <= 0 || (int64_t)filt_n_mel * filt_n_fft > (1 << 20))
        return fail("bad mel filter block");
    filters.resize((size_t)filt_n_mel * filt_n_fft);
    if (!rd(f, filters.data(), filters.size() * 4)) return fail("truncated mel filters");
    int32_t n_tok = 0;
    if (!rd(f, &n_tok, 4) || n_tok < 0 || n_tok > 200000) return fail("bad vocabulary size");
    tokens.resize(n_tok);
    for (int i = 0; i < n_tok; i++) {
        uint32_t len = 0;
        if (!rd(f, &len, 4) || len > 4096) return fail("bad token length");
        tokens[i].resize(len);
        if (len && !rd(f, &tokens[i][0], len)) return fail("truncated vocabulary");
    }
    if (!read_tensor_index(f, err)) { fclose(f); return false; }
    fclose(f);
    if (tensors.empty()) { *err = path + ": no tensors"; return false; }
    return true;
}

// The tensor index shared by every legacy-ggml container (whisper models and the Silero VAD file): from the current position to EOF.
bool GgmlFile::read_tensor_index(FILE* f, std::string* err) {
    auto fail = [&](const std::string& m) { *err = path + ": " + m; return false; };
    int64_t file_size = 0;
    {   // every tensor is bounded by the file's real size before anything is multiplied or sought
        const int64_t here = ftello(f);
        if (here < 0 || fseeko(f, 0, SEEK_END) != 0) return fail("cannot seek");
        file_size = ftello(f);
        if (fseeko(f, here, SEEK_SET) != 0) return fail("cannot seek");
    }
    while (true) {
        int32_t n_dims = 0, name_len = 0, type = 0;
        if (!rd(f, &n_dims, 4)) break;  // EOF
        if (!rd(f, &name_len, 4) || !rd(f, &type, 4)) return fail("truncated tensor header");
        if (n_dims < 1 || n_dims > 4 || name_len <= 0 || name_len > 256) return fail("bad tensor header");
        GgmlTensorInfo t;
        t.type = type;
        t.n_elem = 1;
        for (int i = 0; i < n_dims; i++) {
            int32_t ne = 0;
            if (!rd(f, &ne, 4) || ne <= 0) return fail("bad tensor shape");
            t.ne.push_back(ne);
            if (t.n_elem > file_size / ne) return fail("tensor shape larger than the file");  // also keeps the running product from overflowing
            t.n_elem *= ne;
        }
        std::string name(name_len, 0);
        if (!rd(f, &name[0], name_len)) return fail("truncated tensor name");
        if (type != 0 && type != 1) return fail("tensor '" + name + "' is quantised (type " + std::to_string(type) + "): only f32 / f16 checkpoints are supported");
        t.file_offset = ftello(f);
        const int64_t bytes = t.n_elem * (type == 0 ? 4 : 2);
        if (t.file_offset < 0 || bytes <= 0 || bytes > file_size - t.file_offset) return fail("tensor '" + name + "' runs past the end of the file");
        if (fseeko(f, (off_t)bytes, SEEK_CUR) != 0) return fail("truncated tensor data");
        tensors[name] = t;
    }
    return true;
}

// ggml-silero-v5.1.2.bin (whisper.cpp whisper_vad_init_from_file_with_params; written by models/convert-silero-vad-to-ggml.py):
//   u32 magic | i32 len, model type string ("silero-16k") | i32 version major, minor, patch | i32 n_encoder_layers,
//   {i32 in_channels, out_channels, kernel_size} per layer | i32 lstm_input_size, lstm_hidden_size, final_conv_in, final_conv_out |
//   tensors until EOF (same records as the whisper container).
bool GgmlFile::open_silero(const char* p, SileroHeader* h, std::string* err) {
    path = p ? p : "";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open " + path; return false; }
    auto fail = [&](const std::string& m) { *err = path + ": " + m; fclose(f); return false; };
    uint32_t magic = 0;
    if (!rd(f, &magic, 4) || magic != 0x67676d6cu) return fail("not a ggml file (bad magic)");
    int32_t len = 0;
    if (!rd(f, &len, 4) || len <= 0 || len > 256) return fail("bad model-type string");
    h->model_type.assign((size_t)len, 0);
    if (!rd(f, &h->model_type[0], (size_t)len)) return fail("truncated model-type string");
    if (!rd(f, h->version, 12)) return fail("truncated version");
    if (!rd(f, &h->n_encoder_layers, 4) || h->n_encoder_layers <= 0 || h->n_encoder_layers > 16) return fail("bad encoder layer count");
    h->enc_in.resize(h->n_encoder_layers); h->enc_out.resize(h->n_encoder_layers); h->enc_kernel.resize(h->n_encoder_layers);
    for (int i = 0; i < h->n_encoder_layers; i++)
        if (!rd(f, &h->enc_in[i], 4) || !rd(f, &h->enc_out[i], 4) || !rd(f, &h->enc_kernel[i], 4)) return fail("truncated encoder hyper-parameters");
    if (!rd(f, &h->lstm_input, 4) || !rd(f, &h->lstm_hidden, 4) || !rd(f, &h->final_in, 4) || !rd(f, &h->final_out, 4)) return fail("truncated hyper-parameters");
    if (!read_tensor_index(f, err)) { fclose(f); return false; }
    fclose(f);
    if (tensors.empty()) { *err = path + ": no tensors"; return false; }
    return true;
}

bool GgmlFile::read_f32(const std::string& name, int64_t expect_elems, std::vector<float>* out, std::string* err) const {
    auto it = tensors.find(name);
    if (it == tensors.end()) { *err = path + ": tensor '" + name + "' is missing"; return false; }
    const GgmlTensorInfo& t = it->second;
    if (t.n_elem != expect_elems) { *err = path + ": tensor '" + name + "' has " + std::to_string(t.n_elem) + " elements, expected " + std::to_string(expect_elems); return false; }
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot reopen " + path; return false; }
    bool ok = fseeko(f, (off_t)t.file_offset, SEEK_SET) == 0;
    out->resize((size_t)t.n_elem);
    if (ok && t.type == 0) ok = rd(f, out->data(), (size_t)t.n_elem * 4);
    else if (ok) {
        std::vector<uint16_t> h((size_t)t.n_elem);
        ok = rd(f, h.data(), h.size() * 2);
        if (ok) for (size_t i = 0; i < h.size(); i++) (*out)[i] = f16_to_f32(h[i]);
    }
    fclose(f);
    if (!ok) *err = path + ": short read of tensor '" + name + "'";
    return ok;
}

}  // namespace wdr
