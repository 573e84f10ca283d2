// tokenizer.cu — whisper_tokenize: text -> token ids with the context's vocabulary.  Host code only.
//
// The crate feeds the text of the previous speech segment back as `initial_prompt` (reference src/transcribe.rs:383-386, 502;
// `set_initial_prompt`, :74-76 for the user's own prompt), and whisper_full turns that string into prompt tokens with
// whisper.cpp's tokenizer before anything else.  Restated from whisper.cpp `tokenize()` [UPSTREAM-RECALL]: the text is cut into
// words by the GPT-2 style pattern below (std::regex, ECMAScript, "C" locale: the character classes are ASCII), and every word is
// covered from the left by the longest vocabulary entry that matches at the current position; a byte no entry covers is
// skipped.  A later duplicate string in the vocabulary overrides an earlier one (token_to_id is a map filled in id order).
#include <regex>
#include <string>
#include <unordered_map>
#include <vector>
#include "common.cuh"
#include "model.cuh"
#include "vocab.cuh"

namespace wdr {

static const char* kWordPattern = R"('s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s[:alpha:][:digit:]]+|\s+(?!\S)|\s+)";

int tokenize_text(const std::unordered_map<std::string, int32_t>& token_to_id, const char* text, std::vector<int32_t>& out) {
    out.clear();
    std::vector<std::string> words;
    {
        std::string str = text;
        const std::regex re(kWordPattern);
        std::smatch m;
        while (std::regex_search(str, m, re)) {
            for (auto x : m) words.push_back(x);
            str = m.suffix();
        }
    }
    for (const std::string& word : words) {
        if (word.empty()) continue;
        const int n = (int)word.size();
        int i = 0;
        while (i < n) {
            int j = n;
            bool found = false;
            while (j > i) {
                auto it = token_to_id.find(word.substr(i, j - i));
                if (it != token_to_id.end()) {
                    out.push_back(it->second);
                    i = j;
                    found = true;
                    break;
                }
                --j;
            }
            if (!found) ++i;  // whisper.cpp logs "unknown token" and moves on
        }
    }
    return (int)out.size();
}

// token string -> id over the whole vocabulary of the context (checkpoint strings where it has them, bracketed names for the
// special ids, as whisper.cpp fills token_to_id)
std::unordered_map<std::string, int32_t> context_token_map(const wdr_context* ctx) {
    Vocab v = make_vocab(ctx->arch.n_vocab);
    if (!ctx->file_tokens.empty()) {
        v.file_tokens = &ctx->file_tokens;
        for (int i = 0; i < (int)ctx->file_tokens.size() && i < v.eot; i++)
            if (ctx->file_tokens[i] == " ") { v.space = i; break; }
    }
    std::unordered_map<std::string, int32_t> m;
    m.reserve((size_t)v.n_vocab * 2);
    for (int i = 0; i < v.n_vocab; i++) m[token_text(v, i)] = i;
    return m;
}

int tokenize_for_context(const wdr_context* ctx, const char* text, std::vector<int32_t>& out) {
    // the map is rebuilt per call (51 k short strings, a few ms): prompts are tokenised once per whisper_full call
    return tokenize_text(context_token_map(ctx), text, out);
}

}  // namespace wdr

using namespace wdr;

static int emit_tokens(const std::vector<int32_t>& toks, int32_t* tokens, int n_max_tokens) {
    if ((int)toks.size() > n_max_tokens) return -(int)toks.size();  // whisper_tokenize: the negated count when the buffer is too small
    for (size_t i = 0; i < toks.size(); i++) tokens[i] = toks[i];
    return (int)toks.size();
}

extern "C" int wdr_tokenize(wdr_context* ctx, const char* text, int32_t* tokens, int n_max_tokens) {
    clear_error();
    if (!ctx || !text || n_max_tokens < 0 || (n_max_tokens > 0 && !tokens)) { set_error("wdr_tokenize: bad arguments"); return 0; }
    std::vector<int32_t> toks;
    tokenize_for_context(ctx, text, toks);
    return emit_tokens(toks, tokens, n_max_tokens);
}

extern "C" int wdr_tokenize_with_vocab(const char* const* token_strings, int n_tokens, const char* text, int32_t* tokens, int n_max_tokens) {
    clear_error();
    if (!token_strings || n_tokens < 0 || !text || n_max_tokens < 0 || (n_max_tokens > 0 && !tokens)) { set_error("wdr_tokenize_with_vocab: bad arguments"); return 0; }
    std::unordered_map<std::string, int32_t> m;
    for (int i = 0; i < n_tokens; i++)
        if (token_strings[i]) m[token_strings[i]] = i;
    std::vector<int32_t> toks;
    tokenize_text(m, text, toks);
    return emit_tokens(toks, tokens, n_max_tokens);
}
