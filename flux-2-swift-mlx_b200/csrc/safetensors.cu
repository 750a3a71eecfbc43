// safetensors.cu — safetensors reader / writer and the pre-quantized checkpoint of the reference, inside the C library
// (SURVEY.md §8f-2). Host code only.
//
// Reference: Loading/PrequantizedCheckpoint.swift — format id "flux2-mlx-prequantized-v1" (:41), metadata keys (:244-257),
// atomic save (:259-270), validation order of load (:290-387): payload integrity from the header's data_offsets (:107-141),
// metadata, key set in both directions against the post-quantization manifest, shapes + dtype categories — all BEFORE the
// destination is touched. The tensors are stored under the framework's own flattened module keys (:8-11), which are the keys
// flux2b_set_tensor takes, so a file written by `flux2 export-quantized` loads here and a file written here loads there.
// The safetensors container itself: 8-byte little-endian header length, a JSON header {"__metadata__": {...},
// "<key>": {"dtype": "F16", "shape": [..], "data_offsets": [begin, end]}, ...}, then the raw little-endian payload.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>

#include "ctx.h"

namespace f2b {

namespace {

struct StEntry {
  std::string dtype;
  std::vector<int64_t> shape;
  uint64_t begin = 0, end = 0;
};
struct StHeader {
  std::map<std::string, StEntry> tensors;
  std::map<std::string, std::string> metadata;
  uint64_t data_start = 0;  // file offset of the payload
  uint64_t payload = 0;     // max data_offsets end
};

// ---- a JSON reader just large enough for safetensors headers (objects, arrays, strings, integers, literals)
struct Json {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  bool str(std::string* out) {
    ws();
    if (p >= e || *p != '"') return ok = false;
    ++p;
    out->clear();
    while (p < e && *p != '"') {
      if (*p == '\\') {
        if (++p >= e) return ok = false;
        switch (*p) {
          case 'n': out->push_back('\n'); break;
          case 't': out->push_back('\t'); break;
          case 'r': out->push_back('\r'); break;
          case 'b': out->push_back('\b'); break;
          case 'f': out->push_back('\f'); break;
          case 'u': {  // \uXXXX: keep ASCII, replace the rest (keys / metadata of this format are ASCII)
            if (e - p < 5) return ok = false;
            unsigned v = 0;
            for (int i = 1; i <= 4; ++i) {
              const char h = p[i];
              v = v * 16 + (h >= '0' && h <= '9' ? h - '0' : h >= 'a' && h <= 'f' ? h - 'a' + 10 : h >= 'A' && h <= 'F' ? h - 'A' + 10 : 0);
            }
            out->push_back(v < 128 ? (char)v : '?');
            p += 4;
            break;
          }
          default: out->push_back(*p);
        }
        ++p;
      } else {
        out->push_back(*p++);
      }
    }
    if (p >= e) return ok = false;
    ++p;
    return true;
  }
  bool integer(int64_t* out) {
    ws();
    const char* s = p;
    if (p < e && (*p == '-' || *p == '+')) ++p;
    int64_t v = 0;
    bool any = false;
    while (p < e && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; }
    if (!any) return ok = false;
    *out = (*s == '-') ? -v : v;
    return true;
  }
  // skip any value
  bool skip() {
    ws();
    if (p >= e) return ok = false;
    if (*p == '"') { std::string s; return str(&s); }
    if (*p == '{' || *p == '[') {
      const char close = *p == '{' ? '}' : ']';
      ++p;
      ws();
      if (eat(close)) return true;
      do {
        if (close == '}') { std::string k; if (!str(&k) || !eat(':')) return ok = false; }
        if (!skip()) return false;
      } while (eat(','));
      return eat(close) ? true : (ok = false);
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
    return true;
  }
};

bool parse_header(const char* json, size_t len, StHeader* h, std::string* err) {
  Json j{json, json + len};
  if (!j.eat('{')) { *err = "header is not a JSON object"; return false; }
  if (j.eat('}')) return true;
  do {
    std::string key;
    if (!j.str(&key) || !j.eat(':')) { *err = "malformed header key"; return false; }
    if (key == "__metadata__") {
      if (!j.eat('{')) { *err = "malformed __metadata__"; return false; }
      if (!j.eat('}')) {
        do {
          std::string k, v;
          if (!j.str(&k) || !j.eat(':') || !j.str(&v)) { *err = "malformed __metadata__ entry"; return false; }
          h->metadata[k] = v;
        } while (j.eat(','));
        if (!j.eat('}')) { *err = "malformed __metadata__"; return false; }
      }
      continue;
    }
    StEntry t;
    bool have_off = false;
    if (!j.eat('{')) { *err = "malformed entry for " + key; return false; }
    do {
      std::string f;
      if (!j.str(&f) || !j.eat(':')) { *err = "malformed entry for " + key; return false; }
      if (f == "dtype") {
        if (!j.str(&t.dtype)) { *err = "malformed dtype for " + key; return false; }
      } else if (f == "shape") {
        if (!j.eat('[')) { *err = "malformed shape for " + key; return false; }
        if (!j.eat(']')) {
          do { int64_t v; if (!j.integer(&v) || v < 0) { *err = "malformed shape for " + key; return false; } t.shape.push_back(v); } while (j.eat(','));
          if (!j.eat(']')) { *err = "malformed shape for " + key; return false; }
        }
      } else if (f == "data_offsets") {
        int64_t a, b;
        if (!j.eat('[') || !j.integer(&a) || !j.eat(',') || !j.integer(&b) || !j.eat(']') || a < 0 || b < a) { *err = "malformed data_offsets for " + key; return false; }
        t.begin = (uint64_t)a; t.end = (uint64_t)b; have_off = true;
      } else if (!j.skip()) { *err = "malformed entry for " + key; return false; }
    } while (j.eat(','));
    if (!j.eat('}') || !have_off || t.dtype.empty()) { *err = "incomplete entry for " + key; return false; }
    h->payload = std::max(h->payload, t.end);
    h->tensors[key] = std::move(t);
  } while (j.eat(','));
  if (!j.eat('}') || !j.ok) { *err = "malformed header"; return false; }
  return true;
}

int st_dtype(const std::string& s) {
  if (s == "F32") return FLUX2B_F32;
  if (s == "F16") return FLUX2B_F16;
  if (s == "BF16") return FLUX2B_BF16_T;
  if (s == "U32") return FLUX2B_U32;
  if (s == "U8") return FLUX2B_U8;
  if (s == "I32") return FLUX2B_I32;
  return -1;
}
const char* st_name(int dt) {
  switch (dt) {
    case FLUX2B_F32: return "F32";
    case FLUX2B_F16: return "F16";
    case FLUX2B_BF16_T: return "BF16";
    case FLUX2B_U32: return "U32";
    case FLUX2B_U8: return "U8";
    default: return "I32";
  }
}

// A memory-mapped safetensors file with a validated header.
struct StFile {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t size = 0;
  StHeader h;
  ~StFile() {
    if (map) munmap(const_cast<uint8_t*>(map), size);
    if (fd >= 0) close(fd);
  }
  // returns 0, or 1 with *err set (absent / unreadable / truncated / malformed)
  int open_file(const char* path, std::string* err) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) { *err = std::string("cannot open ") + path; return 1; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { *err = std::string("not a safetensors file: ") + path; return 1; }
    size = (size_t)st.st_size;
    map = static_cast<const uint8_t*>(mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0));
    if (map == MAP_FAILED) { map = nullptr; *err = std::string("mmap failed: ") + path; return 1; }
    madvise(const_cast<uint8_t*>(map), size, MADV_SEQUENTIAL);
    uint64_t hlen = 0;
    memcpy(&hlen, map, 8);  // little-endian per spec (x86-64 / aarch64 hosts)
    if (hlen == 0 || hlen >= (512ull << 20) || 8 + hlen > size) { *err = std::string("bad header length: ") + path; return 1; }
    if (!parse_header(reinterpret_cast<const char*>(map + 8), (size_t)hlen, &h, err)) { *err += std::string(": ") + path; return 1; }
    h.data_start = 8 + hlen;
    // payload integrity (PrequantizedCheckpoint.swift:107-141): the file must be exactly header + declared payload
    if (h.data_start + h.payload != size) {
      *err = "payload is incomplete (" + std::to_string(size) + " bytes on disk, header declares " + std::to_string(h.data_start + h.payload) +
             ") - the file is truncated or corrupt: " + path;
      return 1;
    }
    for (auto& kv : h.tensors) {
      const int dt = st_dtype(kv.second.dtype);
      int64_t n = 1;
      for (auto s : kv.second.shape) n *= s;
      if (dt < 0) { *err = "unsupported dtype " + kv.second.dtype + " for " + kv.first; return 1; }
      if ((uint64_t)n * dtype_size(dt) != kv.second.end - kv.second.begin) { *err = "data_offsets do not match shape x dtype for " + kv.first; return 1; }
    }
    return 0;
  }
  const void* data(const StEntry& t) const { return map + h.data_start + t.begin; }
};

std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
    else if (c == '\n') o += "\\n";
    else if ((unsigned char)c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
    else o.push_back(c);
  }
  return o;
}

struct ManifestEntry { std::vector<int64_t> shape; int cat; };  // cat: 0 float, 1 U32 exact, 2 U8 exact
// The post-quantization parameter manifest of Flux2Transformer2DModel for this context's configuration
// (what PrequantizedCheckpoint.swift:344-356 derives from a structure clone): every Linear as a quantized triplet,
// every QK RMSNorm weight as a float vector.
std::map<std::string, ManifestEntry> dit_manifest(const flux2b_ctx* c) {
  std::map<std::string, ManifestEntry> m;
  const flux2b_dit_config& g = c->dit;
  const int64_t D = (int64_t)g.num_attention_heads * g.attention_head_dim;
  const int64_t Hm = (int64_t)((float)D * g.mlp_ratio);
  int bits = 16, group = 64, has_b = 0, sdt = 0;
  quant_params(c->quant, &bits, &group, &has_b, &sdt);
  auto lin = [&](const std::string& base, int64_t out, int64_t in) {
    m[base + ".weight"] = {{out, in * bits / 32}, 1};
    m[base + ".scales"] = {{out, in / group}, has_b ? 0 : 2};
    if (has_b) m[base + ".biases"] = {{out, in / group}, 0};
  };
  auto norm = [&](const std::string& key) { m[key] = {{(int64_t)g.attention_head_dim}, 0}; };
  lin("xEmbedder", D, g.in_channels);
  lin("contextEmbedder", D, g.joint_attention_dim);
  lin("timeGuidanceEmbed.timestepEmbedder.linear1", D, 256);
  lin("timeGuidanceEmbed.timestepEmbedder.linear2", D, D);
  if (g.guidance_embeds) {
    lin("timeGuidanceEmbed.guidanceEmbedder.linear1", D, 256);
    lin("timeGuidanceEmbed.guidanceEmbedder.linear2", D, D);
  }
  lin("doubleStreamModulationImg.linear", 6 * D, D);
  lin("doubleStreamModulationTxt.linear", 6 * D, D);
  lin("singleStreamModulation.linear", 3 * D, D);
  lin("normOut.linear", 2 * D, D);
  lin("projOut", g.out_channels, D);
  for (int i = 0; i < g.num_layers; ++i) {
    const std::string p = "transformerBlocks." + std::to_string(i) + ".";
    for (const char* n : {"attn.toQ", "attn.toK", "attn.toV", "attn.addQProj", "attn.addKProj", "attn.addVProj", "attn.toOut", "attn.toAddOut"})
      lin(p + n, D, D);
    for (const char* ff : {"ff", "ffContext"}) {
      lin(p + ff + ".activation.proj", 2 * Hm, D);
      lin(p + ff + ".linearOut", D, Hm);
    }
    for (const char* n : {"attn.normQ.weight", "attn.normK.weight", "attn.normAddedQ.weight", "attn.normAddedK.weight"}) norm(p + n);
  }
  for (int i = 0; i < g.num_single_layers; ++i) {
    const std::string p = "singleTransformerBlocks." + std::to_string(i) + ".";
    lin(p + "attn.toQkvMlp", 3 * D + 2 * Hm, D);
    lin(p + "attn.toOut", D, D + Hm);
    norm(p + "attn.normQ.weight");
    norm(p + "attn.normK.weight");
  }
  return m;
}

// Shared metadata validation (PrequantizedCheckpoint.swift:170-202). Returns "" when everything matches.
std::string check_metadata(std::map<std::string, std::string>& md, int quant, const char* source_name, const char* source_fingerprint);

const char* quant_name(int q) {
  static const char* n[] = {"bf16", "qint8", "int4", "mxfp8", "mxfp4", "nvfp4"};
  return (q >= 0 && q <= 5) ? n[q] : "?";
}
const char* mode_name(int q) { return q == 3 ? "mxfp8" : q == 4 ? "mxfp4" : q == 5 ? "nvfp4" : "affine"; }

std::string check_metadata(std::map<std::string, std::string>& md, int quant, const char* source_name, const char* source_fingerprint) {
  int bits = 0, group = 0;
  if (!quant_params(quant, &bits, &group, nullptr, nullptr)) return "quantization is bf16: nothing pre-quantized";
  const std::pair<const char*, std::string> expected[] = {
      {"format", "flux2-mlx-prequantized-v1"}, {"quantization", quant_name(quant)}, {"bits", std::to_string(bits)},
      {"group_size", std::to_string(group)}, {"mode", mode_name(quant)}, {"component", "transformer"}};
  for (auto& kv : expected) {
    auto it = md.find(kv.first);
    if (it == md.end() || it->second != kv.second)
      return std::string("metadata mismatch (") + kv.first + ": " + (it == md.end() ? "nil" : it->second) + " != " + kv.second + ")";
  }
  if (source_name && md["source"] != source_name) return "metadata mismatch (source: " + md["source"] + " != " + source_name + ")";
  if (source_fingerprint && md["source_fingerprint"] != source_fingerprint)
    return "stale: the source weights changed since the export (source_fingerprint)";
  return "";
}

}  // namespace
}  // namespace f2b

using namespace f2b;

extern "C" {

int flux2b_load_safetensors(flux2b_ctx* c, const char* path) {
  if (!c || !path) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  StFile f;
  std::string err;
  if (f.open_file(path, &err)) return fail(FLUX2B_ERR_WEIGHT_LOADING, err);
  int n = 0;
  for (auto& kv : f.h.tensors) {
    const StEntry& t = kv.second;
    std::vector<int64_t> shape = t.shape.empty() ? std::vector<int64_t>{1} : t.shape;
    F2B_TRY(flux2b_set_tensor(c, kv.first.c_str(), f.data(t), st_dtype(t.dtype), shape.data(), (int)shape.size()));
    ++n;
  }
  return n;
}

int flux2b_save_prequantized(flux2b_ctx* c, const char* path, const char* source_name, const char* source_fingerprint, int lora_baked) {
  if (!c || !path) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  if (!c->has_dit || !c->finalized) return fail(FLUX2B_ERR_MODEL_NOT_LOADED, "transformer weights not finalized");
  if (c->quant == FLUX2B_BF16)  // PrequantizedCheckpoint.swift:234-237
    return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "Pre-quantized export requires a quantized model (got bf16). Use the original checkpoint instead.");
  cudaSetDevice(c->device);
  // the full flattened parameter set of the transformer, in key order; it must be exactly the manifest
  const auto manifest = dit_manifest(c);
  std::vector<std::pair<std::string, const Tensor*>> items;
  Tensor ones;  // an RMSNorm weight that was never handed over holds the module default (ones): written out so that the
                // reference's key-set check (PrequantizedCheckpoint.swift:358-365) accepts the file
  for (auto& kv : manifest) {
    auto it = c->tensors.find(kv.first);
    if (it == c->tensors.end()) {
      if (!(kv.second.cat == 0 && kv.second.shape.size() == 1))
        return fail(FLUX2B_ERR_WEIGHT_LOADING, "cannot export: tensor missing from the context: " + kv.first);
      if (!ones.buf.p) {
        ones.dtype = FLUX2B_F16; ones.shape = kv.second.shape;
        std::vector<uint16_t> h((size_t)ones.numel(), 0x3C00);
        if (ones.buf.alloc(h.size() * 2) != cudaSuccess || cudaMemcpy(ones.buf.p, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
          return fail(FLUX2B_ERR_INSUFFICIENT_MEMORY, "export scratch");
      }
      items.emplace_back(kv.first, &ones);
      continue;
    }
    items.emplace_back(kv.first, &it->second);
  }
  int bits = 0, group = 0;
  quant_params(c->quant, &bits, &group, nullptr, nullptr);
  std::string hdr = "{\"__metadata__\":{";
  const std::pair<const char*, std::string> meta[] = {
      {"format", "flux2-mlx-prequantized-v1"}, {"quantization", quant_name(c->quant)}, {"bits", std::to_string(bits)},
      {"group_size", std::to_string(group)}, {"mode", mode_name(c->quant)}, {"component", "transformer"},
      {"source", source_name ? source_name : ""}, {"source_fingerprint", source_fingerprint ? source_fingerprint : "unknown"},
      {"created_by", "flux2b (B200)"}};
  bool first = true;
  for (auto& kv : meta) {
    hdr += std::string(first ? "" : ",") + "\"" + kv.first + "\":\"" + json_escape(kv.second) + "\"";
    first = false;
  }
  if (lora_baked) hdr += ",\"lora_baked\":\"true\"";
  hdr += "}";
  uint64_t off = 0;
  for (auto& it : items) {
    const Tensor& t = *it.second;
    const uint64_t bytes = (uint64_t)t.numel() * dtype_size(t.dtype);
    hdr += ",\"" + json_escape(it.first) + "\":{\"dtype\":\"" + st_name(t.dtype) + "\",\"shape\":[";
    for (size_t i = 0; i < t.shape.size(); ++i) hdr += (i ? "," : "") + std::to_string(t.shape[i]);
    hdr += "],\"data_offsets\":[" + std::to_string(off) + "," + std::to_string(off + bytes) + "]}";
    off += bytes;
  }
  hdr += "}";
  while (hdr.size() % 8) hdr.push_back(' ');
  // atomic write: temporary file in the destination directory, then one rename (PrequantizedCheckpoint.swift:259-270)
  const std::string p(path);
  const size_t slash = p.find_last_of('/');
  const std::string tmp = (slash == std::string::npos ? std::string() : p.substr(0, slash + 1)) + ".tmp-" + (slash == std::string::npos ? p : p.substr(slash + 1));
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) return fail(FLUX2B_ERR_WEIGHT_LOADING, "cannot create " + tmp);
  bool ok = true;
  const uint64_t hlen = hdr.size();
  ok = ok && fwrite(&hlen, 8, 1, fp) == 1 && fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
  std::vector<uint8_t> host;
  for (auto& it : items) {
    if (!ok) break;
    const Tensor& t = *it.second;
    const size_t bytes = (size_t)t.numel() * dtype_size(t.dtype);
    host.resize(bytes);
    if (cudaMemcpy(host.data(), t.buf.p, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
    ok = fwrite(host.data(), 1, bytes, fp) == bytes;
  }
  ok = ok && fflush(fp) == 0 && fsync(fileno(fp)) == 0;
  ok = (fclose(fp) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path) != 0) {
    remove(tmp.c_str());
    return fail(FLUX2B_ERR_WEIGHT_LOADING, std::string("writing the pre-quantized checkpoint failed (previous file, if any, left untouched): ") + path);
  }
  return 0;
}

int flux2b_load_prequantized(flux2b_ctx* c, const char* path, const char* source_name, const char* source_fingerprint) {
  if (!c || !path) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "null argument");
  if (!c->has_dit) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "context has no transformer configuration");
  if (c->quant == FLUX2B_BF16) return fail(FLUX2B_ERR_INVALID_CONFIGURATION, "context quantization is bf16: nothing pre-quantized to load");
  auto skip = [&](const std::string& why) { set_error("pre-quantized checkpoint not applied, fall back to the standard load: " + why); return 1; };
  // 0. payload integrity, header
  StFile f;
  std::string err;
  if (f.open_file(path, &err)) return skip(err);
  // 1. metadata
  const std::string why = check_metadata(f.h.metadata, c->quant, source_name, source_fingerprint);
  if (!why.empty()) return skip(why);
  // 2. key sets in both directions, 3. shapes and dtype categories — before anything is touched (:338-372)
  const auto manifest = dit_manifest(c);
  size_t missing = 0, extra = 0;
  std::string ex_missing, ex_extra;
  for (auto& kv : manifest)
    if (!f.h.tensors.count(kv.first)) { if (!missing++) ex_missing = kv.first; }
  for (auto& kv : f.h.tensors)
    if (!manifest.count(kv.first)) { if (!extra++) ex_extra = kv.first; }
  if (missing || extra)
    return skip("key set mismatch (missing " + std::to_string(missing) + ", extra " + std::to_string(extra) + "; e.g. " + ex_missing + " / " + ex_extra + ")");
  for (auto& kv : manifest) {
    const StEntry& t = f.h.tensors[kv.first];
    const int dt = st_dtype(t.dtype);
    const bool is_float = dt == FLUX2B_F32 || dt == FLUX2B_F16 || dt == FLUX2B_BF16_T;
    const bool dt_ok = kv.second.cat == 0 ? is_float : kv.second.cat == 1 ? dt == FLUX2B_U32 : dt == FLUX2B_U8;
    if (t.shape != kv.second.shape || !dt_ok) return skip("tensor mismatch at " + kv.first + " (shape / dtype " + t.dtype + ")");
  }
  // all checks passed: hand the tensors over (packed layers are taken as they are, finalize skips the quantize pass)
  for (auto& kv : f.h.tensors) {
    const StEntry& t = kv.second;
    F2B_TRY(flux2b_set_tensor(c, kv.first.c_str(), f.data(t), st_dtype(t.dtype), t.shape.data(), (int)t.shape.size()));
  }
  // (:322-325) a LoRA-baked export restyles every generation of that model / quant: say so, loudly, but load it
  if (f.h.metadata.count("lora_baked") && f.h.metadata["lora_baked"] == "true")
    set_error("warning: pre-quantized checkpoint has LoRA weights BAKED IN - every generation with it carries that LoRA");
  else
    set_error("");
  return 0;
}

/* Flux2PrequantizedCheckpoint.isValid (PrequantizedCheckpoint.swift:150-166): payload integrity + header + metadata only, no
 * tensor is read and no device is needed. 1 = valid for this quantization / source, 0 = absent, truncated, foreign or stale
 * (reason in flux2b_last_error()). */
int flux2b_prequantized_is_valid(const char* path, int quant, const char* source_name, const char* source_fingerprint) {
  if (!path) return 0;
  StFile f;
  std::string err;
  if (f.open_file(path, &err)) { set_error(err); return 0; }
  const std::string why = check_metadata(f.h.metadata, quant, source_name, source_fingerprint);
  if (!why.empty()) { set_error(why); return 0; }
  return 1;
}

}  // extern "C"
