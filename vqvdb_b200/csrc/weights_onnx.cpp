// ONNX-initializer reader: the reference's `OnnxModelPaths` / model-directory sources
// (src/core/IVQVAECodec.hpp:27-33, src/backends/onnx/OnnxBackendFactory.cpp:97-145) without ONNX Runtime.
//
// Only the weights are taken from the two graphs (encoder.onnx, decoder.onnx as written by python/to_onnx.py,
// opset 11); the arithmetic is this library's own.  A ~100-line protobuf wire walker reads
//   ModelProto.graph(7) -> GraphProto.initializer(5) = TensorProto{dims(1), data_type(2), float_data(4), name(8), raw_data(9)}
//                       -> GraphProto.node(1)        = NodeProto{input(1), output(2), name(3), op_type(4)}
// and maps tensors back to state_dict names:
//   * conv weights / biases and the codebook keep their names behind a "vqvae." prefix;
//   * GroupNorm is exported as InstanceNormalization + Mul + Add with ANONYMOUS [C,1,1,1] initializers: the Mul / Add
//     node's scope ("/encoder/pre/3/gn1/Mul_2") names the module -> "encoder.pre.3.gn1.weight" / ".bias";
//   * the bias-free Linear layers of the channel attention become MatMul with an anonymous TRANSPOSED [in,out]
//     initializer under "/encoder/attn/fc/0/MatMul" -> "encoder.attn.fc.0.weight" [out,in].
// The result is serialised as a VQVDBW01 pack, so everything downstream is the one code path of weights.cpp.
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "weights.hpp"

namespace vqvdb {

namespace {

struct Span {
	const unsigned char* p = nullptr;
	size_t n = 0;
};

struct Reader {
	const unsigned char* p;
	const unsigned char* end;
	explicit Reader(Span s) : p(s.p), end(s.p + s.n) {}
	bool done() const { return p >= end; }
	uint64_t varint() {
		uint64_t r = 0;
		for (int shift = 0; shift < 64; shift += 7) {
			if (p >= end) throw std::runtime_error("onnx: truncated varint");
			const unsigned char c = *p++;
			r |= (uint64_t)(c & 0x7f) << shift;
			if (!(c & 0x80)) return r;
		}
		throw std::runtime_error("onnx: varint too long");
	}
	// next field: returns false at the end; for wire type 2 `bytes` is the payload, for 0 `value` the varint
	bool next(uint32_t& field, uint32_t& wire, uint64_t& value, Span& bytes) {
		if (done()) return false;
		const uint64_t key = varint();
		field = (uint32_t)(key >> 3);
		wire = (uint32_t)(key & 7);
		value = 0;
		bytes = Span{};
		switch (wire) {
			case 0: value = varint(); break;
			case 1: skip(8); break;
			case 5: skip(4); break;
			case 2: {
				const uint64_t len = varint();
				if (len > (uint64_t)(end - p)) throw std::runtime_error("onnx: truncated field");
				bytes = Span{p, (size_t)len};
				p += len;
				break;
			}
			default: throw std::runtime_error("onnx: unsupported wire type");
		}
		return true;
	}
	void skip(size_t n) {
		if (n > (size_t)(end - p)) throw std::runtime_error("onnx: truncated field");
		p += n;
	}
};

struct OnnxTensor {
	std::string name;
	std::vector<int> dims;
	std::vector<float> data;
};

OnnxTensor parse_tensor(Span s) {
	OnnxTensor t;
	Reader r(s);
	uint32_t f, w;
	uint64_t v;
	Span b;
	int dtype = 0;
	Span raw{};
	std::vector<float> fdata;
	while (r.next(f, w, v, b)) {
		if (f == 1) {  // dims: repeated int64, packed or not
			if (w == 0) t.dims.push_back((int)v);
			else {
				Reader pr(b);
				while (!pr.done()) t.dims.push_back((int)pr.varint());
			}
		} else if (f == 2) dtype = (int)v;
		else if (f == 8) t.name.assign(reinterpret_cast<const char*>(b.p), b.n);
		else if (f == 9) raw = b;
		else if (f == 4 && w == 2) {  // packed float_data
			fdata.resize(b.n / 4);
			std::memcpy(fdata.data(), b.p, fdata.size() * 4);
		}
	}
	if (dtype != 1) return t;  // only FLOAT tensors carry weights; others (int64 shape constants) are left empty
	size_t numel = 1;
	for (int d : t.dims) numel *= (size_t)d;
	if (raw.n == numel * 4) {
		t.data.resize(numel);
		std::memcpy(t.data.data(), raw.p, raw.n);  // little-endian fp32, as the pack
	} else if (fdata.size() == numel) t.data = std::move(fdata);
	else throw std::runtime_error("onnx: initializer " + t.name + " has no usable float payload");
	return t;
}

struct OnnxNode {
	std::string name, op;
	std::vector<std::string> in, out;
};

OnnxNode parse_node(Span s) {
	OnnxNode n;
	Reader r(s);
	uint32_t f, w;
	uint64_t v;
	Span b;
	while (r.next(f, w, v, b)) {
		if (w != 2) continue;
		std::string str(reinterpret_cast<const char*>(b.p), b.n);
		if (f == 1) n.in.push_back(std::move(str));
		else if (f == 2) n.out.push_back(std::move(str));
		else if (f == 3) n.name = std::move(str);
		else if (f == 4) n.op = std::move(str);
	}
	return n;
}

// "/encoder/pre/3/gn1/Mul_2" -> "encoder.pre.3.gn1"
std::string module_of(const OnnxNode& n) {
	std::string s = !n.name.empty() ? n.name : (n.out.empty() ? std::string() : n.out[0]);
	const size_t cut = s.find_last_of('/');
	if (cut == std::string::npos) return std::string();
	s = s.substr(0, cut);
	while (!s.empty() && s[0] == '/') s.erase(0, 1);
	for (char& c : s)
		if (c == '/') c = '.';
	return s;
}

void read_graph(const std::vector<unsigned char>& model, std::vector<std::pair<std::string, OnnxTensor>>& out) {
	Reader mr(Span{model.data(), model.size()});
	uint32_t f, w;
	uint64_t v;
	Span b, graph{};
	while (mr.next(f, w, v, b))
		if (f == 7 && w == 2) graph = b;
	if (!graph.p) throw std::runtime_error("onnx: no graph in model");
	std::vector<OnnxTensor> inits;
	std::vector<OnnxNode> nodes;
	Reader gr(graph);
	while (gr.next(f, w, v, b)) {
		if (w != 2) continue;
		if (f == 5) inits.push_back(parse_tensor(b));
		else if (f == 1) nodes.push_back(parse_node(b));
	}
	auto find_init = [&](const std::string& name) -> OnnxTensor* {
		for (auto& t : inits)
			if (t.name == name && !t.data.empty()) return &t;
		return nullptr;
	};
	for (auto& t : inits) {
		if (t.data.empty()) continue;
		if (t.name.rfind("vqvae.", 0) == 0) out.emplace_back(t.name.substr(6), t);
	}
	for (const auto& n : nodes) {
		if (n.op != "Mul" && n.op != "Add" && n.op != "MatMul") continue;
		for (const auto& in : n.in) {
			if (in.rfind("onnx::", 0) != 0) continue;
			OnnxTensor* t = find_init(in);
			const std::string mod = module_of(n);
			if (!t || mod.empty()) continue;  // the top-level MatMul is the (transposed) codebook: already taken by name
			OnnxTensor r;
			if (n.op == "MatMul") {
				if (t->dims.size() != 2) continue;
				const int ki = t->dims[0], ko = t->dims[1];  // [in, out] -> Linear.weight [out, in]
				r.dims = {ko, ki};
				r.data.resize(t->data.size());
				for (int i = 0; i < ki; ++i)
					for (int o = 0; o < ko; ++o) r.data[(size_t)o * ki + i] = t->data[(size_t)i * ko + o];
				out.emplace_back(mod + ".weight", std::move(r));
			} else {
				r.dims = {t->dims.empty() ? 1 : t->dims[0]};
				if ((size_t)r.dims[0] != t->data.size()) continue;  // not a per-channel affine vector
				r.data = t->data;
				out.emplace_back(mod + (n.op == "Mul" ? ".weight" : ".bias"), std::move(r));
			}
		}
	}
}

void put_u32(std::vector<unsigned char>& v, uint32_t x) {
	for (int i = 0; i < 4; ++i) v.push_back((unsigned char)(x >> (8 * i)));
}
void put_u64(std::vector<unsigned char>& v, uint64_t x) {
	for (int i = 0; i < 8; ++i) v.push_back((unsigned char)(x >> (8 * i)));
}

}  // namespace

std::vector<unsigned char> read_file_bytes(const std::string& path) {
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f) throw std::runtime_error("cannot open " + path);
	const std::streamsize sz = f.tellg();
	f.seekg(0);
	std::vector<unsigned char> buf((size_t)sz);
	if (sz > 0 && !f.read(reinterpret_cast<char*>(buf.data()), sz)) throw std::runtime_error("cannot read " + path);
	return buf;
}

std::vector<unsigned char> onnx_to_pack(const std::vector<unsigned char>& encoder_onnx, const std::vector<unsigned char>& decoder_onnx) {
	std::vector<std::pair<std::string, OnnxTensor>> all, tensors;
	read_graph(encoder_onnx, all);
	read_graph(decoder_onnx, all);
	for (auto& kv : all) {  // the codebook is in both graphs: keep the first of each name
		bool dup = false;
		for (auto& have : tensors) dup = dup || have.first == kv.first;
		if (!dup) tensors.push_back(std::move(kv));
	}
	const OnnxTensor *pre = nullptr, *emb = nullptr;
	for (auto& kv : tensors) {
		if (kv.first == "encoder.pre.0.weight") pre = &kv.second;
		if (kv.first == "quantizer.embedding") emb = &kv.second;
	}
	if (!pre || pre->dims.size() != 5 || !emb || emb->dims.size() != 2) throw std::runtime_error("onnx: encoder.pre.0.weight / quantizer.embedding not found");
	// VQVDBW01 (tools/weights_pack.py): magic, n, in_channels, D, K, entries {name, dims, offset, nbytes}, payload size, 64-byte aligned payload
	std::vector<unsigned char> head(8);
	std::memcpy(head.data(), "VQVDBW01", 8);
	put_u32(head, (uint32_t)tensors.size());
	put_u32(head, (uint32_t)pre->dims[1]);
	put_u32(head, (uint32_t)emb->dims[1]);
	put_u32(head, (uint32_t)emb->dims[0]);
	uint64_t off = 0;
	std::vector<uint64_t> offs;
	for (auto& kv : tensors) {
		put_u32(head, (uint32_t)kv.first.size());
		head.insert(head.end(), kv.first.begin(), kv.first.end());
		put_u32(head, (uint32_t)kv.second.dims.size());
		for (int d : kv.second.dims) put_u32(head, (uint32_t)d);
		const uint64_t nb = kv.second.data.size() * 4;
		put_u64(head, off);
		put_u64(head, nb);
		offs.push_back(off);
		off = (off + nb + 63) & ~uint64_t(63);
	}
	put_u64(head, off);
	head.resize((head.size() + 63) & ~size_t(63), 0);
	std::vector<unsigned char> blob(head.size() + (size_t)off, 0);
	std::memcpy(blob.data(), head.data(), head.size());
	for (size_t i = 0; i < tensors.size(); ++i)
		std::memcpy(blob.data() + head.size() + offs[i], tensors[i].second.data.data(), tensors[i].second.data.size() * 4);
	return blob;
}

}  // namespace vqvdb
