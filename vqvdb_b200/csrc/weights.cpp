#include "weights.hpp"

#include <cstring>
#include <fstream>
#include <stdexcept>

namespace vqvdb {

namespace {
struct Cursor {
	const unsigned char* p;
	size_t size, pos = 0;
	void need(size_t n) const {
		if (pos + n > size) throw std::runtime_error("weight pack truncated");
	}
	uint32_t u32() {
		need(4);
		uint32_t v;
		std::memcpy(&v, p + pos, 4);
		pos += 4;
		return v;
	}
	uint64_t u64() {
		need(8);
		uint64_t v;
		std::memcpy(&v, p + pos, 8);
		pos += 8;
		return v;
	}
};
}  // namespace

void WeightPack::parse(const void* data, size_t size) {
	if (!data || size < 32) throw std::runtime_error("weight pack is empty");
	blob.assign(static_cast<const unsigned char*>(data), static_cast<const unsigned char*>(data) + size);
	Cursor c{blob.data(), blob.size()};
	if (std::memcmp(blob.data(), "VQVDBW01", 8) != 0) throw std::runtime_error("not a VQVDBW01 weight pack");
	c.pos = 8;
	const uint32_t n = c.u32();
	in_channels = (int)c.u32();
	embedding_dim = (int)c.u32();
	num_embeddings = (int)c.u32();
	if (n == 0 || n > 4096) throw std::runtime_error("weight pack: implausible tensor count");
	struct Entry {
		std::string name;
		std::vector<int> dims;
		uint64_t off, nbytes;
	};
	std::vector<Entry> entries(n);
	for (auto& e : entries) {
		const uint32_t ln = c.u32();
		c.need(ln);
		e.name.assign(reinterpret_cast<const char*>(blob.data() + c.pos), ln);
		c.pos += ln;
		const uint32_t nd = c.u32();
		if (nd > 8) throw std::runtime_error("weight pack: tensor rank too large");
		size_t numel = 1;
		for (uint32_t d = 0; d < nd; ++d) {
			e.dims.push_back((int)c.u32());
			numel *= (size_t)e.dims.back();
		}
		e.off = c.u64();
		e.nbytes = c.u64();
		if (e.nbytes != numel * sizeof(float)) throw std::runtime_error("weight pack: size mismatch for " + e.name);
	}
	const uint64_t payload = c.u64();
	const size_t base = (c.pos + 63) & ~size_t(63);
	if (base + payload != blob.size()) throw std::runtime_error("weight pack: payload size mismatch");
	tensors.clear();
	for (auto& e : entries) {
		if (e.off + e.nbytes > payload || (e.off & 3)) throw std::runtime_error("weight pack: bad offset for " + e.name);
		PackTensor t;
		t.dims = e.dims;
		t.data = reinterpret_cast<const float*>(blob.data() + base + e.off);
		tensors.emplace(e.name, std::move(t));
	}
}

void WeightPack::load_file(const std::string& path) {
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f) throw std::runtime_error("weight pack not found: " + path);
	const std::streamsize sz = f.tellg();
	f.seekg(0);
	std::vector<unsigned char> buf((size_t)sz);
	if (!f.read(reinterpret_cast<char*>(buf.data()), sz)) throw std::runtime_error("cannot read weight pack: " + path);
	parse(buf.data(), buf.size());
}

const PackTensor& WeightPack::get(const std::string& name) const {
	auto it = tensors.find(name);
	if (it == tensors.end()) throw std::runtime_error("weight pack: missing tensor " + name);
	return it->second;
}

std::vector<float> transpose_conv_weight(const PackTensor& w) {
	if (w.dims.size() != 5) throw std::runtime_error("conv weight must be rank 5");
	const int co = w.dims[0], ci = w.dims[1], k3 = w.dims[2] * w.dims[3] * w.dims[4];
	std::vector<float> out((size_t)co * ci * k3);
	for (int o = 0; o < co; ++o)
		for (int i = 0; i < ci; ++i)
			for (int t = 0; t < k3; ++t) out[((size_t)i * k3 + t) * co + o] = w.data[((size_t)o * ci + i) * k3 + t];
	return out;
}

}  // namespace vqvdb
