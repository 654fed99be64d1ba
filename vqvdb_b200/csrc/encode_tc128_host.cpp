// Host-side preparation of the tensor-core vec3 encoder's weight streams and parameter blocks (encode_tc128.cuh).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>

#include "encode_tc128_stream.hpp"
#include "fp16_split.hpp"

namespace vqvdb {

namespace {

constexpr size_t kUnitBytes = kEnc128UnitBytes;

bool dims_are(const WeightPack& p, const std::string& name, std::initializer_list<int> dims) {
	const auto it = p.tensors.find(name);
	return it != p.tensors.end() && it->second.dims == std::vector<int>(dims);
}

// One [64 n][64 k] tile of a unit: element (n, k) at byte n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2.
// w is [cout][cin_total][27]; the tile takes output channels oc0.., input channels ic0.., filter tap `tap`.
void fill_tile(uint8_t* tile, const float* w, int cin_total, int oc0, int ic0, int tap, int part) {
	for (int n = 0; n < 64; ++n)
		for (int k = 0; k < 64; ++k) {
			const uint16_t b = split_part(w[((size_t)(oc0 + n) * cin_total + (ic0 + k)) * 27 + tap], part);
			const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
			std::memcpy(tile + off, &b, 2);
		}
}

// The 18 steps of one pass of a 128 -> 128 conv on the 4^3 grid: 9 (kd, kh) pairs x 2 input-channel halves, each the hi
// unit then the lo unit, a unit = [3 kw] tiles of output channels oc0 .. oc0 + 63.
uint8_t* fill_pass128(uint8_t* u, const float* w, int oc0) {
	for (int pair = 0; pair < 9; ++pair)
		for (int khalf = 0; khalf < 2; ++khalf)
			for (int part = 0; part < 2; ++part, u += kUnitBytes)
				for (int kw = 0; kw < 3; ++kw) fill_tile(u + (size_t)kw * 8192, w, 128, oc0, khalf * 64, pair * 3 + kw, part);
	return u;
}

// One [128 n][64 k] tile (the stride-2 conv: all 128 output channels, one filter tap), same row layout.
void fill_tile_down(uint8_t* tile, const float* w, int tap, int part) {
	for (int n = 0; n < 128; ++n)
		for (int k = 0; k < 64; ++k) {
			const uint16_t b = split_part(w[((size_t)n * 64 + k) * 27 + tap], part);
			const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
			std::memcpy(tile + off, &b, 2);
		}
}

// M[k][c] = sum_d e_kd W_dc in double
std::vector<double> fold_proj_into_codebook(const WeightPack& p) {
	const PackTensor& e = p.get("quantizer.embedding");  // [256][128]
	const PackTensor& w = p.get("encoder.proj.weight");  // [128 d][128 c][1][1][1]
	std::vector<double> m((size_t)256 * 128, 0.0);
	for (int k = 0; k < 256; ++k)
		for (int d = 0; d < 128; ++d) {
			const double ev = e.data[(size_t)k * 128 + d];
			for (int c = 0; c < 128; ++c) m[(size_t)k * 128 + c] += ev * (double)w.data[(size_t)d * 128 + c];
		}
	return m;
}

}  // namespace

bool encoder128_supports(const WeightPack& p) {
	if (p.embedding_dim != 128 || p.num_embeddings != 256 || p.in_channels != 3) return false;
	bool ok = dims_are(p, "quantizer.embedding", {256, 128}) && dims_are(p, "encoder.pre.0.weight", {64, 3, 3, 3, 3}) &&
	          dims_are(p, "encoder.pre.0.bias", {64}) && dims_are(p, "encoder.pre.1.weight", {64}) && dims_are(p, "encoder.pre.1.bias", {64}) &&
	          dims_are(p, "encoder.down1.weight", {128, 64, 3, 3, 3}) && dims_are(p, "encoder.down1.bias", {128}) &&
	          dims_are(p, "encoder.attn.fc.0.weight", {32, 128}) && dims_are(p, "encoder.attn.fc.2.weight", {128, 32}) &&
	          dims_are(p, "encoder.proj.weight", {128, 128, 1, 1, 1}) && dims_are(p, "encoder.proj.bias", {128}) &&
	          p.tensors.find("encoder.res_stack.2.conv1.weight") == p.tensors.end();
	for (const char* v : {".gn1.weight", ".gn1.bias", ".gn2.weight", ".gn2.bias", ".conv1.bias", ".conv2.bias"})
		ok = ok && dims_are(p, std::string("encoder.pre.3") + v, {64});
	ok = ok && dims_are(p, "encoder.pre.3.conv1.weight", {64, 64, 3, 3, 3}) && dims_are(p, "encoder.pre.3.conv2.weight", {64, 64, 3, 3, 3});
	for (int r = 0; r < 2 && ok; ++r) {
		const std::string pre = "encoder.res_stack." + std::to_string(r);
		for (const char* v : {".gn1.weight", ".gn1.bias", ".gn2.weight", ".gn2.bias", ".conv1.bias", ".conv2.bias"}) ok = ok && dims_are(p, pre + v, {128});
		ok = ok && dims_are(p, pre + ".conv1.weight", {128, 128, 3, 3, 3}) && dims_are(p, pre + ".conv2.weight", {128, 128, 3, 3, 3});
	}
	return ok;
}

std::vector<uint8_t> build_encoder128_back_units(const WeightPack& p) {
	if (!encoder128_supports(p)) throw std::runtime_error("encoder128 units: unsupported architecture");
	std::vector<uint8_t> out((size_t)kEnc128BackUnits * kUnitBytes);
	uint8_t* u = out.data();
	for (const char* name : {"encoder.res_stack.0.conv1.weight", "encoder.res_stack.0.conv2.weight", "encoder.res_stack.1.conv1.weight",
	                         "encoder.res_stack.1.conv2.weight"}) {
		const float* w = p.get(name).data;
		for (int h = 0; h < 2; ++h) u = fill_pass128(u, w, h * 64);
	}
	if (u != out.data() + out.size()) throw std::logic_error("encoder128 back unit stream size mismatch");
	return out;
}

std::vector<float> build_encoder128_back_params(const WeightPack& p) {
	if (!encoder128_supports(p)) throw std::runtime_error("encoder128 params: unsupported architecture");
	std::vector<float> out;
	auto put = [&](const std::string& name, size_t n) {
		const PackTensor& t = p.get(name);
		if (t.numel() != n) throw std::runtime_error("encoder128 params: unexpected size of " + name);
		out.insert(out.end(), t.data, t.data + n);
	};
	for (int r = 0; r < 2; ++r) {
		const std::string pre = "encoder.res_stack." + std::to_string(r);
		for (const char* v : {".gn1.weight", ".gn1.bias", ".conv1.bias", ".gn2.weight", ".gn2.bias", ".conv2.bias"}) put(pre + v, 128);
	}
	put("encoder.proj.bias", 128);
	// the codebook search's constants (encode_tc128.cu): per-code offsets of the folded scores and what bounds their error
	{
		const PackTensor& e = p.get("quantizer.embedding");
		const PackTensor& w = p.get("encoder.proj.weight");
		const PackTensor& b = p.get("encoder.proj.bias");
		const std::vector<double> m = fold_proj_into_codebook(p);
		double m_max = 0.0, e_max = 0.0, b2 = 0.0;
		for (int k = 0; k < 256; ++k) {
			double e2 = 0.0, be = 0.0, m2 = 0.0;
			for (int d = 0; d < 128; ++d) {
				e2 += (double)e.data[(size_t)k * 128 + d] * e.data[(size_t)k * 128 + d];
				be += (double)e.data[(size_t)k * 128 + d] * b.data[d];
			}
			for (int c = 0; c < 128; ++c) m2 += m[(size_t)k * 128 + c] * m[(size_t)k * 128 + c];
			out.push_back((float)(e2 - 2.0 * be));
			m_max = std::max(m_max, std::sqrt(m2));
			e_max = std::max(e_max, std::sqrt(e2));
		}
		// |W|_2 <= min(|W|_F, sqrt(|W|_1 |W|_inf))
		double fro2 = 0.0, n1 = 0.0, ninf = 0.0;
		std::vector<double> col(128, 0.0);
		for (int d = 0; d < 128; ++d) {
			double rowsum = 0.0;
			for (int c = 0; c < 128; ++c) {
				const double v = std::fabs((double)w.data[(size_t)d * 128 + c]);
				fro2 += v * v;
				rowsum += v;
				col[c] += v;
			}
			ninf = std::max(ninf, rowsum);
		}
		for (int c = 0; c < 128; ++c) n1 = std::max(n1, col[c]);
		for (int d = 0; d < 128; ++d) b2 += (double)b.data[d] * b.data[d];
		const double w_norm = std::min(std::sqrt(fro2), std::sqrt(n1 * ninf));
		out.push_back(std::nextafter((float)m_max, INFINITY));
		out.push_back(std::nextafter((float)w_norm, INFINITY));
		out.push_back(std::nextafter((float)(std::sqrt(b2) + e_max), INFINITY));
		out.push_back(0.f);
	}
	if (out.size() != (size_t)par128e::total) throw std::logic_error("encoder128 back parameter block size mismatch");
	return out;
}

std::vector<uint8_t> build_encoder128_front_units(const WeightPack& p) {
	if (!encoder128_supports(p)) throw std::runtime_error("encoder128 units: unsupported architecture");
	std::vector<uint8_t> out((size_t)kEnc128FrontUnits * kUnitBytes, 0);
	uint8_t* u = out.data();
	for (const char* name : {"encoder.pre.3.conv1.weight", "encoder.pre.3.conv2.weight"}) {
		const float* w = p.get(name).data;
		for (int pair = 0; pair < 9; ++pair)
			for (int part = 0; part < 2; ++part, u += kUnitBytes)
				for (int kw = 0; kw < 3; ++kw) fill_tile(u + (size_t)kw * 8192, w, 64, 0, 0, pair * 3 + kw, part);
	}
	const float* dw = p.get("encoder.down1.weight").data;
	for (int tap = 0; tap < 27; ++tap)
		for (int part = 0; part < 2; ++part, u += kUnitBytes) fill_tile_down(u, dw, tap, part);
	if (u != out.data() + out.size()) throw std::logic_error("encoder128 front unit stream size mismatch");
	return out;
}

std::vector<float> build_encoder128_front_params(const WeightPack& p) {
	if (!encoder128_supports(p)) throw std::runtime_error("encoder128 params: unsupported architecture");
	std::vector<float> out;
	auto put = [&](const std::string& name, size_t n) {
		const PackTensor& t = p.get(name);
		if (t.numel() != n) throw std::runtime_error("encoder128 params: unexpected size of " + name);
		out.insert(out.end(), t.data, t.data + n);
	};
	put("encoder.pre.0.bias", 64);
	put("encoder.pre.1.weight", 64);
	put("encoder.pre.1.bias", 64);
	for (const char* v : {".gn1.weight", ".gn1.bias", ".conv1.bias", ".gn2.weight", ".gn2.bias", ".conv2.bias"}) put(std::string("encoder.pre.3") + v, 64);
	put("encoder.down1.bias", 128);
	if (out.size() != (size_t)par128f::total) throw std::logic_error("encoder128 front parameter block size mismatch");
	return out;
}

std::vector<float> build_encoder128_vq_fold(const WeightPack& p) {
	if (!encoder128_supports(p)) throw std::runtime_error("encoder128 VQ fold: unsupported architecture");
	const std::vector<double> m = fold_proj_into_codebook(p);
	return std::vector<float>(m.begin(), m.end());
}

std::vector<uint8_t> build_encoder128_vq_units(const WeightPack& p) {
	const std::vector<float> m = build_encoder128_vq_fold(p);
	std::vector<uint8_t> out((size_t)kEnc128VqUnits * kEnc128VqUnitBytes);
	uint8_t* u = out.data();
	for (int khalf = 0; khalf < 2; ++khalf)
		for (int nhalf = 0; nhalf < 2; ++nhalf)
			for (int part = 0; part < 2; ++part, u += kEnc128VqUnitBytes)
				for (int n = 0; n < 128; ++n)
					for (int k = 0; k < 64; ++k) {
						const uint16_t b = split_part(m[(size_t)(nhalf * 128 + n) * 128 + khalf * 64 + k], part);
						const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
						std::memcpy(u + off, &b, 2);
					}
	return out;
}

}  // namespace vqvdb
