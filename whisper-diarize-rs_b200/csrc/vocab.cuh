// vocab.cuh — special-token ids as whisper.cpp derives them from n_vocab and the synthetic token
// strings (no tokenizer file can exist offline; the test tree restates the same mapping for the checker).
// Special tokens use whisper.cpp's bracket names, which the reference's control-token filter relies on
// (reference src/transcribe.rs:206-212).
#pragma once
#include <string>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace wdr {

struct Vocab {
    int n_vocab, eot, sot, translate, transcribe, solm, prev, nosp, not_, beg, lang0, n_langs, space;
    bool multilingual;
    const std::vector<std::string>* file_tokens = nullptr;  // token strings of a ggml checkpoint (ids below its size)
};

inline Vocab make_vocab(int n_vocab) {
    Vocab v;
    v.n_vocab = n_vocab;
    v.multilingual = n_vocab >= 51865;
    v.eot = 50256 + (v.multilingual ? 1 : 0);
    v.sot = v.eot + 1;
    v.beg = n_vocab - 1501;
    v.translate = v.beg - 6; v.transcribe = v.beg - 5; v.solm = v.beg - 4; v.prev = v.beg - 3; v.nosp = v.beg - 2; v.not_ = v.beg - 1;
    v.lang0 = v.sot + 1;
    v.n_langs = 100;
    v.space = 220;
    return v;
}

static const char* const kLangs[100] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi", "he", "uk",
    "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn",
    "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn",
    "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my",
    "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su", "yue"};

inline int lang_id_from_str(const char* s) {
    if (!s) return -1;
    for (int i = 0; i < 100; i++)
        if (strcmp(kLangs[i], s) == 0) return i;
    return -1;
}

inline std::string token_text(const Vocab& v, int i) {
    char buf[48];
    if (v.file_tokens && i >= 0 && i < (int)v.file_tokens->size() && i < v.eot) return (*v.file_tokens)[i];
    if (i < v.eot) {
        if (i == v.space) return " ";
        std::string s;
        int n = i;
        do { s.push_back((char)('a' + n % 26)); n /= 26; } while (n > 0);
        std::string out = (i % 3 == 0) ? " " : "";
        out += s;
        if (i % 17 == 0) out += ",";
        else if (i % 29 == 0) out += ".";
        else if (i % 31 == 0) out += "?";
        else if (i % 23 == 0) out.push_back((char)('0' + i % 10));
        return out;
    }
    if (i == v.eot) return "[_EOT_]";
    if (i == v.sot) return "[_SOT_]";
    if (i == v.translate) return "[_TRANSLATE_]";
    if (i == v.transcribe) return "[_TRANSCRIBE_]";
    if (i == v.solm) return "[_SOLM_]";
    if (i == v.prev) return "[_PREV_]";
    if (i == v.nosp) return "[_NOSP_]";
    if (i == v.not_) return "[_NOT_]";
    if (i == v.beg) return "[_BEG_]";
    if (i > v.beg) { snprintf(buf, sizeof(buf), "[_TT_%d]", i - v.beg); return buf; }
    const int n_lang_tokens = v.n_vocab - 51765 - (v.multilingual ? 1 : 0);
    if (i >= v.lang0 && i < v.lang0 + n_lang_tokens) { snprintf(buf, sizeof(buf), "[_LANG_%s]", kLangs[i - v.lang0]); return buf; }
    snprintf(buf, sizeof(buf), "[_extra_token_%d]", i);
    return buf;
}

// whisper.cpp voice_length
inline float voice_length(const std::string& text) {
    float r = 0.0f;
    for (char c : text) {
        if (c == ' ') r += 0.01f;
        else if (c == ',') r += 2.00f;
        else if (c == '.' || c == '!' || c == '?') r += 3.00f;
        else if (c >= '0' && c <= '9') r += 3.00f;
        else r += 1.00f;
    }
    return r;
}

}  // namespace wdr
