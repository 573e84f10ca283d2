// onnx_file.cu — see onnx_file.cuh.  Host-only: protobuf wire-format walk of an ONNX ModelProto, no third-party code.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include "onnx_file.cuh"

namespace wdr {

namespace {

// ---- protobuf wire format ------------------------------------------------------------------------------------------------
struct Cur {
    const unsigned char* p;
    const unsigned char* end;
    bool ok = true;
    Cur(const unsigned char* b, size_t n) : p(b), end(b + n) {}
    bool done() const { return p >= end; }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            const unsigned char c = *p++;
            v |= (uint64_t)(c & 0x7F) << shift;
            if (!(c & 0x80)) return v;
        }
        ok = false;  // more than 10 bytes
        return 0;
    }
    // key -> (field, wire type); false at the end of the buffer or on a malformed key
    bool key(uint32_t* field, int* wt) {
        if (p >= end) return false;
        const uint64_t k = varint();
        if (!ok) return false;
        *field = (uint32_t)(k >> 3);
        *wt = (int)(k & 7);
        return true;
    }
    Cur sub() {  // length-delimited payload
        const uint64_t n = varint();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return Cur(p, 0); }
        Cur c(p, (size_t)n);
        p += n;
        return c;
    }
    std::string str() {
        Cur c = sub();
        return ok ? std::string(reinterpret_cast<const char*>(c.p), (size_t)(c.end - c.p)) : std::string();
    }
    uint32_t fixed32() {
        if (end - p < 4) { ok = false; return 0; }
        uint32_t v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    uint64_t fixed64() {
        if (end - p < 8) { ok = false; return 0; }
        uint64_t v;
        memcpy(&v, p, 8);
        p += 8;
        return v;
    }
    void skip(int wt) {
        switch (wt) {
            case 0: varint(); break;
            case 1: fixed64(); break;
            case 2: sub(); break;
            case 5: fixed32(); break;
            default: ok = false;  // groups (3, 4) do not occur in ONNX
        }
    }
};

float half_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
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

// TensorProto: dims 1, data_type 2, float_data 4, int32_data 5, int64_data 7, name 8, raw_data 9, double_data 10, data_location 14
bool parse_tensor(Cur c, OnnxTensor* t, std::string* err) {
    const unsigned char* raw = nullptr;
    size_t raw_n = 0;
    std::vector<float> fdata;
    std::vector<int64_t> i32data, i64data;
    std::vector<double> ddata;
    int64_t location = 0;
    uint32_t f;
    int wt;
    while (c.key(&f, &wt)) {
        if (f == 1 && wt == 0) t->dims.push_back((int64_t)c.varint());
        else if (f == 1 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { t->dims.push_back((int64_t)s.varint()); if (!s.ok) c.ok = false; } }
        else if (f == 2 && wt == 0) t->dtype = (int)c.varint();
        else if (f == 4 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { uint32_t u = s.fixed32(); if (!s.ok) { c.ok = false; break; } float v; memcpy(&v, &u, 4); fdata.push_back(v); } }
        else if (f == 4 && wt == 5) { uint32_t u = c.fixed32(); float v; memcpy(&v, &u, 4); fdata.push_back(v); }
        else if (f == 5 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { i32data.push_back((int64_t)(int32_t)s.varint()); if (!s.ok) c.ok = false; } }
        else if (f == 5 && wt == 0) i32data.push_back((int64_t)(int32_t)c.varint());
        else if (f == 7 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { i64data.push_back((int64_t)s.varint()); if (!s.ok) c.ok = false; } }
        else if (f == 7 && wt == 0) i64data.push_back((int64_t)c.varint());
        else if (f == 8 && wt == 2) t->name = c.str();
        else if (f == 9 && wt == 2) { Cur s = c.sub(); raw = s.p; raw_n = (size_t)(s.end - s.p); }
        else if (f == 10 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { uint64_t u = s.fixed64(); if (!s.ok) { c.ok = false; break; } double v; memcpy(&v, &u, 8); ddata.push_back(v); } }
        else if (f == 14 && wt == 0) location = (int64_t)c.varint();
        else c.skip(wt);
        if (!c.ok) { *err = "malformed TensorProto '" + t->name + "'"; return false; }
    }
    if (!c.ok) { *err = "malformed TensorProto '" + t->name + "'"; return false; }
    if (location == 1) { *err = "tensor '" + t->name + "' keeps its data in an external file: not supported"; return false; }
    int64_t n = 1;
    for (int64_t d : t->dims) {
        if (d < 0 || (d > 0 && n > (int64_t)1 << 40)) { *err = "tensor '" + t->name + "': implausible shape"; return false; }
        n *= d;
    }
    auto need = [&](size_t have, const char* what) {
        if ((int64_t)have != n) { *err = "tensor '" + t->name + "': " + what + " holds " + std::to_string(have) + " values, shape says " + std::to_string(n); return false; }
        return true;
    };
    switch (t->dtype) {
        case 1:  // FLOAT
            if (raw) { if (!need(raw_n / 4, "raw_data") || raw_n % 4) return false; t->f32.resize((size_t)n); if (n) memcpy(t->f32.data(), raw, (size_t)n * 4); }
            else { if (!need(fdata.size(), "float_data")) return false; t->f32.swap(fdata); }
            break;
        case 10:  // FLOAT16 (raw, or one value per int32_data entry)
            t->f32.resize((size_t)n);
            if (raw) { if (!need(raw_n / 2, "raw_data")) return false; for (int64_t i = 0; i < n; i++) { uint16_t h; memcpy(&h, raw + 2 * i, 2); t->f32[(size_t)i] = half_to_float(h); } }
            else { if (!need(i32data.size(), "int32_data")) return false; for (int64_t i = 0; i < n; i++) t->f32[(size_t)i] = half_to_float((uint16_t)i32data[(size_t)i]); }
            break;
        case 11:  // DOUBLE
            t->f32.resize((size_t)n);
            if (raw) { if (!need(raw_n / 8, "raw_data")) return false; for (int64_t i = 0; i < n; i++) { double v; memcpy(&v, raw + 8 * i, 8); t->f32[(size_t)i] = (float)v; } }
            else { if (!need(ddata.size(), "double_data")) return false; for (int64_t i = 0; i < n; i++) t->f32[(size_t)i] = (float)ddata[(size_t)i]; }
            break;
        case 7:  // INT64
            if (raw) { if (!need(raw_n / 8, "raw_data")) return false; t->i64.resize((size_t)n); if (n) memcpy(t->i64.data(), raw, (size_t)n * 8); }
            else { if (!need(i64data.size(), "int64_data")) return false; t->i64.swap(i64data); }
            break;
        case 6:  // INT32
            t->i64.resize((size_t)n);
            if (raw) { if (!need(raw_n / 4, "raw_data")) return false; for (int64_t i = 0; i < n; i++) { int32_t v; memcpy(&v, raw + 4 * i, 4); t->i64[(size_t)i] = v; } }
            else { if (!need(i32data.size(), "int32_data")) return false; t->i64.swap(i32data); }
            break;
        default: break;  // other element types never carry the parameters looked for here: shape only
    }
    return true;
}

// AttributeProto: name 1, f 2, i 3, s 4, t 5, floats 7, ints 8
bool parse_attr(Cur c, OnnxAttr* a, std::string* err) {
    uint32_t f;
    int wt;
    while (c.key(&f, &wt)) {
        if (f == 1 && wt == 2) a->name = c.str();
        else if (f == 2 && wt == 5) { uint32_t u = c.fixed32(); memcpy(&a->f, &u, 4); }
        else if (f == 3 && wt == 0) a->i = (int64_t)c.varint();
        else if (f == 4 && wt == 2) a->s = c.str();
        else if (f == 5 && wt == 2) { Cur s = c.sub(); if (c.ok) { a->has_tensor = true; if (!parse_tensor(s, &a->t, err)) return false; } }
        else if (f == 7 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { uint32_t u = s.fixed32(); if (!s.ok) { c.ok = false; break; } float v; memcpy(&v, &u, 4); a->floats.push_back(v); } }
        else if (f == 7 && wt == 5) { uint32_t u = c.fixed32(); float v; memcpy(&v, &u, 4); a->floats.push_back(v); }
        else if (f == 8 && wt == 2) { Cur s = c.sub(); while (c.ok && !s.done()) { a->ints.push_back((int64_t)s.varint()); if (!s.ok) c.ok = false; } }
        else if (f == 8 && wt == 0) a->ints.push_back((int64_t)c.varint());
        else c.skip(wt);
        if (!c.ok) { *err = "malformed AttributeProto"; return false; }
    }
    if (!c.ok) { *err = "malformed AttributeProto"; return false; }
    return true;
}

// NodeProto: input 1, output 2, name 3, op_type 4, attribute 5
bool parse_node(Cur c, OnnxNode* n, std::string* err) {
    uint32_t f;
    int wt;
    while (c.key(&f, &wt)) {
        if (f == 1 && wt == 2) n->inputs.push_back(c.str());
        else if (f == 2 && wt == 2) n->outputs.push_back(c.str());
        else if (f == 3 && wt == 2) n->name = c.str();
        else if (f == 4 && wt == 2) n->op_type = c.str();
        else if (f == 5 && wt == 2) { Cur s = c.sub(); if (c.ok) { n->attrs.emplace_back(); if (!parse_attr(s, &n->attrs.back(), err)) return false; } }
        else c.skip(wt);
        if (!c.ok) { *err = "malformed NodeProto"; return false; }
    }
    if (!c.ok) { *err = "malformed NodeProto"; return false; }
    return true;
}

std::string value_info_name(Cur c) {  // ValueInfoProto: name 1
    uint32_t f;
    int wt;
    std::string name;
    while (c.key(&f, &wt)) {
        if (f == 1 && wt == 2) name = c.str();
        else c.skip(wt);
        if (!c.ok) break;
    }
    return name;
}

}  // namespace

const OnnxAttr* OnnxNode::attr(const char* n) const {
    for (auto& a : attrs)
        if (a.name == n) return &a;
    return nullptr;
}
int64_t OnnxNode::attr_i(const char* n, int64_t dflt) const { const OnnxAttr* a = attr(n); return a ? a->i : dflt; }
float OnnxNode::attr_f(const char* n, float dflt) const { const OnnxAttr* a = attr(n); return a ? a->f : dflt; }

const OnnxTensor* OnnxFile::tensor(const std::string& value_name) const {
    auto it = tensors.find(value_name);
    return it == tensors.end() ? nullptr : &it->second;
}

bool OnnxFile::parse(const unsigned char* data, size_t size, std::string* err) {
    nodes.clear(); tensors.clear(); inputs.clear(); outputs.clear();
    Cur m(data, size);
    uint32_t f;
    int wt;
    bool have_graph = false;
    std::vector<std::string> graph_inputs;
    while (m.key(&f, &wt)) {
        if (f == 1 && wt == 0) ir_version = (int64_t)m.varint();
        else if (f == 2 && wt == 2) producer = m.str();
        else if (f == 7 && wt == 2) {  // GraphProto: node 1, initializer 5, input 11, output 12
            Cur g = m.sub();
            if (!m.ok) break;
            have_graph = true;
            uint32_t gf;
            int gw;
            while (g.key(&gf, &gw)) {
                if (gf == 1 && gw == 2) { Cur s = g.sub(); if (g.ok) { nodes.emplace_back(); if (!parse_node(s, &nodes.back(), err)) return false; } }
                else if (gf == 5 && gw == 2) {
                    Cur s = g.sub();
                    if (g.ok) {
                        OnnxTensor t;
                        if (!parse_tensor(s, &t, err)) return false;
                        const std::string name = t.name;
                        tensors[name] = std::move(t);
                    }
                }
                else if (gf == 11 && gw == 2) { Cur s = g.sub(); if (g.ok) graph_inputs.push_back(value_info_name(s)); }
                else if (gf == 12 && gw == 2) { Cur s = g.sub(); if (g.ok) outputs.push_back(value_info_name(s)); }
                else g.skip(gw);
                if (!g.ok) { *err = "malformed GraphProto"; return false; }
            }
            if (!g.ok) { *err = "malformed GraphProto"; return false; }
        } else m.skip(wt);
        if (!m.ok) break;
    }
    if (!m.ok) { *err = "malformed ModelProto (not an ONNX file?)"; return false; }
    if (!have_graph || nodes.empty()) { *err = "no graph in the file (not an ONNX model?)"; return false; }
    for (auto& n : nodes)  // Constant nodes carry parameters too (exporters that do not lift them to initializers)
        if (n.op_type == "Constant" && !n.outputs.empty()) {
            const OnnxAttr* a = n.attr("value");
            if (a && a->has_tensor) { OnnxTensor t = a->t; t.name = n.outputs[0]; tensors[n.outputs[0]] = std::move(t); }
        }
    for (auto& s : graph_inputs)
        if (!tensors.count(s)) inputs.push_back(s);
    return true;
}

bool OnnxFile::load(const char* p, std::string* err) {
    path = p ? p : "";
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) { *err = "cannot open " + path; return false; }
    std::vector<unsigned char> buf;
    bool ok = fseeko(fp, 0, SEEK_END) == 0;
    const int64_t size = ok ? (int64_t)ftello(fp) : -1;
    ok = ok && size > 0 && size < ((int64_t)1 << 31) && fseeko(fp, 0, SEEK_SET) == 0;  // protobuf messages are < 2 GiB
    if (ok) {
        buf.resize((size_t)size);
        ok = fread(buf.data(), 1, (size_t)size, fp) == (size_t)size;
    }
    fclose(fp);
    if (!ok) { *err = path + ": cannot read the file"; return false; }
    std::string e;
    if (!parse(buf.data(), buf.size(), &e)) { *err = path + ": " + e; return false; }
    return true;
}

// ---- parameter extraction --------------------------------------------------------------------------------------------------
namespace {

bool dims_are(const OnnxTensor* t, std::initializer_list<int64_t> want) {
    if (!t || t->dims.size() != want.size() || t->f32.empty()) return false;
    size_t i = 0;
    for (int64_t w : want) if (t->dims[i++] != w) return false;
    return true;
}
std::string dims_str(const OnnxTensor* t) {
    if (!t) return "(not a constant of the graph)";
    std::string s = "[";
    for (size_t i = 0; i < t->dims.size(); i++) s += (i ? "," : "") + std::to_string(t->dims[i]);
    return s + "]";
}

// a Linear layer as the exporters write it: Gemm(x, W, b) or MatMul(x, W^T) followed by Add(., b)
struct LinearOp { std::vector<float> w; std::vector<float> b; int64_t n_out = 0, n_in = 0; };

bool collect_linears(const OnnxFile& f, std::vector<LinearOp>* out, std::string* err) {
    for (size_t i = 0; i < f.nodes.size(); i++) {
        const OnnxNode& n = f.nodes[i];
        if (n.op_type == "Gemm" && n.inputs.size() >= 2) {
            const OnnxTensor* W = f.tensor(n.inputs[1]);
            if (!W || W->dims.size() != 2 || W->f32.empty()) continue;
            if (n.attr_f("alpha", 1.0f) != 1.0f || n.attr_f("beta", 1.0f) != 1.0f || n.attr_i("transA", 0) != 0) { *err = "Gemm '" + n.name + "' uses alpha / beta / transA"; return false; }
            LinearOp l;
            const bool tb = n.attr_i("transB", 0) != 0;
            l.n_out = tb ? W->dims[0] : W->dims[1];
            l.n_in = tb ? W->dims[1] : W->dims[0];
            l.w.resize(W->f32.size());
            for (int64_t o = 0; o < l.n_out; o++)
                for (int64_t k = 0; k < l.n_in; k++) l.w[(size_t)(o * l.n_in + k)] = tb ? W->f32[(size_t)(o * l.n_in + k)] : W->f32[(size_t)(k * l.n_out + o)];
            const OnnxTensor* B = n.inputs.size() >= 3 ? f.tensor(n.inputs[2]) : nullptr;
            l.b.assign((size_t)l.n_out, 0.0f);
            if (B) { if ((int64_t)B->f32.size() != l.n_out) { *err = "Gemm '" + n.name + "': bias length"; return false; } l.b = B->f32; }
            out->push_back(std::move(l));
        } else if (n.op_type == "MatMul" && n.inputs.size() == 2 && !n.outputs.empty()) {
            const OnnxTensor* W = f.tensor(n.inputs[1]);
            if (!W || W->dims.size() != 2 || W->f32.empty()) continue;  // activations x activations (none in these nets)
            LinearOp l;
            l.n_in = W->dims[0];
            l.n_out = W->dims[1];
            l.w.resize(W->f32.size());
            for (int64_t o = 0; o < l.n_out; o++)
                for (int64_t k = 0; k < l.n_in; k++) l.w[(size_t)(o * l.n_in + k)] = W->f32[(size_t)(k * l.n_out + o)];
            l.b.assign((size_t)l.n_out, 0.0f);
            for (size_t j = i + 1; j < f.nodes.size(); j++) {  // the Add that consumes this MatMul's output
                const OnnxNode& a = f.nodes[j];
                if (a.op_type != "Add" || a.inputs.size() != 2) continue;
                const int mine = a.inputs[0] == n.outputs[0] ? 0 : a.inputs[1] == n.outputs[0] ? 1 : -1;
                if (mine < 0) continue;
                const OnnxTensor* B = f.tensor(a.inputs[1 - mine]);
                if (B && (int64_t)B->f32.size() == l.n_out) l.b = B->f32;
                break;
            }
            out->push_back(std::move(l));
        }
    }
    return true;
}

// ONNX LSTM gate order is i, o, f, c; PyTorch's (and this library's) is i, f, g(c), o
void lstm_rows_to_torch(const float* src, int64_t hidden, int64_t cols, std::vector<float>* dst) {
    static const int from[4] = {0, 2, 3, 1};  // torch gate g comes from ONNX gate from[g]
    dst->resize((size_t)(4 * hidden * cols));
    for (int g = 0; g < 4; g++)
        memcpy(dst->data() + (size_t)g * hidden * cols, src + (size_t)from[g] * hidden * cols, sizeof(float) * (size_t)(hidden * cols));
}

}  // namespace

bool onnx_extract_pyannet(const OnnxFile& f, NamedTensors* out, std::string* err) {
    out->clear();
    std::vector<const OnnxNode*> inorm, conv, lstm;
    for (auto& n : f.nodes) {
        if (n.op_type == "InstanceNormalization") inorm.push_back(&n);
        else if (n.op_type == "Conv") conv.push_back(&n);
        else if (n.op_type == "LSTM") lstm.push_back(&n);
    }
    std::vector<LinearOp> lin;
    if (!collect_linears(f, &lin, err)) return false;
    if (inorm.size() != 4 || conv.size() != 3 || lstm.size() != 4 || lin.size() != 3) {
        *err = "not a segmentation-3.0 (PyanNet) export: found " + std::to_string(inorm.size()) + " InstanceNormalization, " + std::to_string(conv.size()) + " Conv, " +
               std::to_string(lstm.size()) + " LSTM, " + std::to_string(lin.size()) + " linear layers; expected 4 / 3 / 4 / 3";
        return false;
    }
    auto put = [&](const std::string& name, const std::vector<float>& v) { (*out)[name] = v; };
    // waveform norm + the three per-stage norms
    static const int64_t norm_c[4] = {1, 80, 60, 60};
    for (int i = 0; i < 4; i++) {
        if (inorm[i]->inputs.size() < 3) { *err = "InstanceNormalization without scale / bias"; return false; }
        const OnnxTensor *g = f.tensor(inorm[i]->inputs[1]), *b = f.tensor(inorm[i]->inputs[2]);
        if (!dims_are(g, {norm_c[i]}) || !dims_are(b, {norm_c[i]})) { *err = "InstanceNormalization #" + std::to_string(i) + ": scale " + dims_str(g) + ", expected [" + std::to_string(norm_c[i]) + "]"; return false; }
        const float eps = inorm[i]->attr_f("epsilon", 1e-5f);
        if (fabsf(eps - 1e-5f) > 1e-9f) { *err = "InstanceNormalization epsilon " + std::to_string(eps) + " (this library computes 1e-5)"; return false; }
        const std::string base = i == 0 ? "wav_norm" : "norm" + std::to_string(i - 1);
        put(base + ".weight", g->f32);
        put(base + ".bias", b->f32);
    }
    static const int64_t conv_d[3][3] = {{80, 1, 251}, {60, 80, 5}, {60, 60, 5}};
    for (int i = 0; i < 3; i++) {
        if (conv[i]->inputs.size() < 2) { *err = "Conv without a weight input"; return false; }
        const OnnxTensor* W = f.tensor(conv[i]->inputs[1]);
        if (!W) { *err = "Conv #" + std::to_string(i) + ": the filters are computed inside the graph (a SincNet filterbank that was not constant-folded at export): not supported"; return false; }
        if (!dims_are(W, {conv_d[i][0], conv_d[i][1], conv_d[i][2]})) { *err = "Conv #" + std::to_string(i) + ": weight " + dims_str(W); return false; }
        put("conv" + std::to_string(i) + ".weight", W->f32);
        const OnnxTensor* B = conv[i]->inputs.size() >= 3 ? f.tensor(conv[i]->inputs[2]) : nullptr;
        if (i == 0) {
            if (B) for (float v : B->f32) if (v != 0.0f) { *err = "the sinc filterbank convolution carries a bias: not a PyanNet front end"; return false; }
        } else {
            if (!dims_are(B, {conv_d[i][0]})) { *err = "Conv #" + std::to_string(i) + ": bias " + dims_str(B); return false; }
            put("conv" + std::to_string(i) + ".bias", B->f32);
        }
    }
    for (int l = 0; l < 4; l++) {
        const OnnxNode* n = lstm[l];
        const int64_t n_in = l == 0 ? 60 : 256;
        if (n->inputs.size() < 4) { *err = "LSTM without W / R / B inputs"; return false; }
        const OnnxAttr* dir = n->attr("direction");
        if (n->attr_i("hidden_size", 0) != 128 || !dir || dir->s != "bidirectional") { *err = "LSTM #" + std::to_string(l) + " is not a bidirectional LSTM(128)"; return false; }
        const OnnxTensor *W = f.tensor(n->inputs[1]), *R = f.tensor(n->inputs[2]), *B = f.tensor(n->inputs[3]);
        if (!dims_are(W, {2, 512, n_in}) || !dims_are(R, {2, 512, 128}) || !dims_are(B, {2, 1024})) {
            *err = "LSTM #" + std::to_string(l) + ": W " + dims_str(W) + " R " + dims_str(R) + " B " + dims_str(B);
            return false;
        }
        for (int d = 0; d < 2; d++) {
            const std::string sfx = "_l" + std::to_string(l) + (d ? "_reverse" : "");
            std::vector<float> t;
            lstm_rows_to_torch(W->f32.data() + (size_t)d * 512 * n_in, 128, n_in, &t);
            put("lstm.weight_ih" + sfx, t);
            lstm_rows_to_torch(R->f32.data() + (size_t)d * 512 * 128, 128, 128, &t);
            put("lstm.weight_hh" + sfx, t);
            lstm_rows_to_torch(B->f32.data() + (size_t)d * 1024, 128, 1, &t);
            put("lstm.bias_ih" + sfx, t);
            lstm_rows_to_torch(B->f32.data() + (size_t)d * 1024 + 512, 128, 1, &t);
            put("lstm.bias_hh" + sfx, t);
        }
    }
    static const int64_t lin_d[3][2] = {{128, 256}, {128, 128}, {7, 128}};
    static const char* lin_n[3] = {"linear0", "linear1", "classifier"};
    for (int i = 0; i < 3; i++) {
        if (lin[i].n_out != lin_d[i][0] || lin[i].n_in != lin_d[i][1]) { *err = std::string(lin_n[i]) + ": " + std::to_string(lin[i].n_out) + " x " + std::to_string(lin[i].n_in); return false; }
        put(std::string(lin_n[i]) + ".weight", lin[i].w);
        put(std::string(lin_n[i]) + ".bias", lin[i].b);
    }
    return true;
}

bool onnx_extract_resnet34(const OnnxFile& f, NamedTensors* out, int* emb_dim, std::string* err) {
    out->clear();
    // the 36 convolutions of wespeaker/models/resnet.py in execution order, a block's shortcut conv after its conv2
    struct Spec { std::string name; int64_t ci, co, k, stride; };
    std::vector<Spec> specs;
    specs.push_back({"conv1", 1, 32, 3, 1});
    {
        static const int64_t planes[4] = {32, 64, 128, 256}, blocks[4] = {3, 4, 6, 3}, strides[4] = {1, 2, 2, 2};
        int64_t c_in = 32;
        for (int li = 0; li < 4; li++)
            for (int64_t bi = 0; bi < blocks[li]; bi++) {
                const int64_t s = bi == 0 ? strides[li] : 1;
                const std::string b = "layer" + std::to_string(li + 1) + "." + std::to_string(bi);
                specs.push_back({b + ".conv1", c_in, planes[li], 3, s});
                specs.push_back({b + ".conv2", planes[li], planes[li], 3, 1});
                if (s != 1 || c_in != planes[li]) specs.push_back({b + ".shortcut", c_in, planes[li], 1, s});
                c_in = planes[li];
            }
    }
    std::vector<const OnnxNode*> conv;
    for (auto& n : f.nodes)
        if (n.op_type == "Conv") conv.push_back(&n);
    if (conv.size() != specs.size()) {
        *err = "not a WeSpeaker ResNet34 export: " + std::to_string(conv.size()) + " Conv nodes, expected " + std::to_string(specs.size()) +
               " (the crate's default download is the CAM++ export, a different network that this library does not run)";
        return false;
    }
    for (size_t i = 0; i < specs.size(); i++) {
        const Spec& sp = specs[i];
        const OnnxNode* n = conv[i];
        const OnnxTensor* W = n->inputs.size() >= 2 ? f.tensor(n->inputs[1]) : nullptr;
        if (!dims_are(W, {sp.co, sp.ci, sp.k, sp.k})) { *err = "Conv #" + std::to_string(i) + " (" + sp.name + "): weight " + dims_str(W); return false; }
        const OnnxAttr* st = n->attr("strides");
        if (st && (st->ints.size() != 2 || st->ints[0] != sp.stride || st->ints[1] != sp.stride)) { *err = "Conv " + sp.name + ": unexpected strides"; return false; }
        std::vector<float> w = W->f32, b((size_t)sp.co, 0.0f);
        const OnnxTensor* B = n->inputs.size() >= 3 ? f.tensor(n->inputs[2]) : nullptr;
        if (B) { if ((int64_t)B->f32.size() != sp.co) { *err = "Conv " + sp.name + ": bias length"; return false; } b = B->f32; }
        // an un-folded export keeps BatchNormalization(conv_out, scale, B, mean, var): fold it (fp32: s = g / sqrt(var + eps), w *= s, b = beta - mean * s)
        for (auto& bn : f.nodes) {
            if (bn.op_type != "BatchNormalization" || bn.inputs.size() < 5 || n->outputs.empty() || bn.inputs[0] != n->outputs[0]) continue;
            const OnnxTensor *g = f.tensor(bn.inputs[1]), *beta = f.tensor(bn.inputs[2]), *mean = f.tensor(bn.inputs[3]), *var = f.tensor(bn.inputs[4]);
            if (!dims_are(g, {sp.co}) || !dims_are(beta, {sp.co}) || !dims_are(mean, {sp.co}) || !dims_are(var, {sp.co})) { *err = "BatchNormalization after " + sp.name + ": parameter shapes"; return false; }
            const float eps = bn.attr_f("epsilon", 1e-5f);
            const int64_t fan = sp.ci * sp.k * sp.k;
            for (int64_t o = 0; o < sp.co; o++) {
                const float s = g->f32[(size_t)o] / sqrtf(var->f32[(size_t)o] + eps);
                for (int64_t j = 0; j < fan; j++) w[(size_t)(o * fan + j)] *= s;
                b[(size_t)o] = beta->f32[(size_t)o] - (mean->f32[(size_t)o] - b[(size_t)o]) * s;
            }
            break;
        }
        (*out)[sp.name + ".weight"] = std::move(w);
        (*out)[sp.name + ".bias"] = std::move(b);
    }
    std::vector<LinearOp> lin;
    if (!collect_linears(f, &lin, err)) return false;
    if (lin.empty() || lin.back().n_in != 5120) { *err = "no embedding layer (Linear 5120 -> D) after the statistics pooling"; return false; }
    *emb_dim = (int)lin.back().n_out;
    (*out)["seg_1.weight"] = lin.back().w;
    (*out)["seg_1.bias"] = lin.back().b;
    return true;
}

}  // namespace wdr
