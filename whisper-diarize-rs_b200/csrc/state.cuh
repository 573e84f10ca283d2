// state.cuh — wdr_state (whisper_state equivalent): stream, workspaces, device-resident results.
#pragma once
#include <string>
#include <vector>
#include "decoder.cuh"
#include "encoder.cuh"
#include "model.cuh"
#include "profile.cuh"

namespace wdr {
// one result segment of the last full call (whisper_segment): times in centiseconds relative to its chunk's first sample
struct ResultSegment {
    int chunk = 0;
    int64_t t0 = 0, t1 = 0;
    std::string text;
    float no_speech_prob = 0.0f;
    std::vector<wdr_token_data> tokens;
    std::vector<std::string> token_text;
};
struct ChunkInfo {
    int seek_delta, failed, completed, n_sampled, has_ts, result_len, seek_end, n_segments;
    float no_speech_prob;
    int lang_id;
    float temperature;  // the temperature whose result stands (0 unless the fallback ladder ran)
};
// staging / scratch of the full pipeline (full.cu)
struct FullScratch {
    void* pcm_dev = nullptr;
    size_t pcm_cap = 0;  // bytes
    int32_t* nvalid_dev = nullptr;
    size_t nvalid_cap = 0;
    float* energy_dev = nullptr;
    size_t energy_cap = 0;
    float* energy_host = nullptr;  // pinned
    size_t energy_host_cap = 0;
    int32_t* done_host = nullptr;  // pinned
    cudaEvent_t ev_energy = nullptr, ev_energy_done = nullptr, ev_h2d = nullptr;
    float* dtw_x = nullptr;
    size_t dtw_x_cap = 0;
    float* dtw_stat = nullptr;
    size_t dtw_stat_cap = 0;
    int32_t* dtw_path = nullptr;
    size_t dtw_path_cap = 0;
    char* dtw_wins = nullptr;  // device DtwWindow table of dtw_run
    size_t dtw_wins_cap = 0;
    int32_t* dtw_path_host = nullptr;  // pinned
    size_t dtw_path_host_cap = 0;
    // the energy pass of A.5 on the device: per-token (t0, t1) in / out, text flags, token counts, thresholds
    int64_t* a5_tok_dev = nullptr;
    size_t a5_tok_cap = 0;
    int64_t* a5_tok_host = nullptr;   // pinned
    size_t a5_tok_host_cap = 0;
    uint8_t* a5_text_dev = nullptr;
    size_t a5_text_cap = 0;
    uint8_t* a5_text_host = nullptr;  // pinned
    size_t a5_text_host_cap = 0;
    int32_t* a5_n_dev = nullptr;
    size_t a5_n_cap = 0;
    int32_t* a5_n_host = nullptr;     // pinned
    size_t a5_n_host_cap = 0;
    float* a5_thold_dev = nullptr;
    size_t a5_thold_cap = 0;
    float* lp_dev = nullptr;       // [rows][ldv] processed log-probabilities of a sampling pass (temperature > 0, greedy strategy)
    size_t lp_dev_cap = 0;
    float* lp_host = nullptr;      // pinned
    size_t lp_host_cap = 0;
    int last_decode_steps = 0;
    // phase boundaries of the last group on the compute stream: start | encoder | cross-KV | greedy decode | DTW pass | DTW
    cudaEvent_t ev_phase[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double phase_ms[5] = {0, 0, 0, 0, 0};  // accumulated over the groups of the last full call
    int decode_steps = 0;                  // greedy iterations of the last full call (summed over groups)
    unsigned long long* cross_stats_host = nullptr;  // pinned [2]
    int64_t cross_launches = 0, cross_live = 0;      // dec_cross_attn_kernel launches / live (launch, window) pairs of the last full call
    void release() {
        cudaFree(pcm_dev); cudaFree(nvalid_dev); cudaFree(energy_dev); cudaFree(dtw_x); cudaFree(dtw_stat); cudaFree(dtw_path); cudaFree(dtw_wins); cudaFree(lp_dev);
        cudaFree(a5_tok_dev); cudaFree(a5_text_dev); cudaFree(a5_n_dev); cudaFree(a5_thold_dev);
        if (a5_tok_host) cudaFreeHost(a5_tok_host);
        if (a5_text_host) cudaFreeHost(a5_text_host);
        if (a5_n_host) cudaFreeHost(a5_n_host);
        if (lp_host) cudaFreeHost(lp_host);
        if (energy_host) cudaFreeHost(energy_host);
        if (dtw_path_host) cudaFreeHost(dtw_path_host);
        if (done_host) cudaFreeHost(done_host);
        if (cross_stats_host) cudaFreeHost(cross_stats_host);
        if (ev_energy) cudaEventDestroy(ev_energy);
        if (ev_energy_done) cudaEventDestroy(ev_energy_done);
        if (ev_h2d) cudaEventDestroy(ev_h2d);
        for (auto e : ev_phase) if (e) cudaEventDestroy(e);
        *this = FullScratch();
    }
};
}  // namespace wdr

struct wdr_state {
    wdr_context* ctx = nullptr;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t copy_stream = nullptr;  // H2D staging, overlapped with compute group by group
    wdr::EncoderWorkspace enc;
    wdr::DecoderWorkspace dec;
    wdr::DtwPassWorkspace dtwp;
    wdr::FullScratch full;
    std::vector<wdr::ResultSegment> results;  // segments of the last full call, chunk order
    std::vector<wdr::ChunkInfo> chunk_info;
    int lang_id = 0;
    wdr::Profiler prof;
    // extra lanes of the full pipeline (full.cu): child states with their own streams / workspaces; lane 0 is this state
    std::vector<wdr_state*> lanes;
    int n_lanes = 0;  // 0 = library default (WDR_LANES, else 1)
    // device-resident staging / results of the last encode call
    int16_t* pcm_dev = nullptr;
    size_t pcm_cap = 0;
    int32_t* nvalid_dev = nullptr;
    int nvalid_cap = 0;
    float* enc_out = nullptr;  // [n_enc_chunks][1500][d] fp32
    size_t enc_out_cap = 0;
    int n_enc_chunks = 0;
    float* digest_dev = nullptr;
    int digest_cap = 0;
    std::vector<cudaEvent_t> copy_events;
};
