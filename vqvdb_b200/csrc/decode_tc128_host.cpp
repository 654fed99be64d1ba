// Host-side preparation of the 128-channel tensor-core decoder's weight stream and parameter block (decode_tc128.cuh).
#include <cstring>
#include <stdexcept>
#include <string>

#include "weights.hpp"

namespace vqvdb {

namespace {

constexpr int kPasses = 13, kUnitsPerPass = 18;
constexpr size_t kUnitBytes = 3 * 8192;

uint16_t f32_to_bf16_rn(float f) {
	uint32_t u;
	std::memcpy(&u, &f, 4);
	if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);  // inf / nan pass through
	u += 0x7fffu + ((u >> 16) & 1u);
	return (uint16_t)(u >> 16);
}

bool dims_are(const WeightPack& p, const std::string& name, std::initializer_list<int> dims) {
	const auto it = p.tensors.find(name);
	return it != p.tensors.end() && it->second.dims == std::vector<int>(dims);
}

// One [64 n][64 k] tile of a unit: element (n, k) at byte n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2.
// w is [cout][cin][27]; the tile takes output channels oc0.., input channels ic0.., filter tap `tap`.
void fill_tile(uint8_t* tile, const float* w, int cin_total, int oc0, int ic0, int tap) {
	for (int n = 0; n < 64; ++n)
		for (int k = 0; k < 64; ++k) {
			const uint16_t b = f32_to_bf16_rn(w[((size_t)(oc0 + n) * cin_total + (ic0 + k)) * 27 + tap]);
			const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
			std::memcpy(tile + off, &b, 2);
		}
}

// The 18 units of one pass: 9 (kd, kh) pairs x 2 input-channel halves, each [3 kw] tiles of output channels oc0 .. oc0+63.
uint8_t* fill_pass(uint8_t* u, const float* w, int oc0) {
	for (int pair = 0; pair < 9; ++pair)
		for (int khalf = 0; khalf < 2; ++khalf, u += kUnitBytes)
			for (int kw = 0; kw < 3; ++kw) fill_tile(u + (size_t)kw * 8192, w, 128, oc0, khalf * 64, pair * 3 + kw);
	return u;
}

}  // namespace

bool decoder128_supports(const WeightPack& p) {
	if (p.embedding_dim != 128 || p.num_embeddings != 256 || p.in_channels != 3) return false;
	bool ok = dims_are(p, "quantizer.embedding", {256, 128}) && dims_are(p, "decoder.stem.0.weight", {128, 128, 3, 3, 3}) &&
	          dims_are(p, "decoder.stem.0.bias", {128}) && dims_are(p, "decoder.stem.1.weight", {128}) && dims_are(p, "decoder.stem.1.bias", {128}) &&
	          dims_are(p, "decoder.attn.fc.0.weight", {32, 128}) && dims_are(p, "decoder.attn.fc.2.weight", {128, 32}) &&
	          dims_are(p, "decoder.up_conv.weight", {256, 128, 3, 3, 3}) && dims_are(p, "decoder.up_conv.bias", {256}) &&
	          dims_are(p, "decoder.final.weight", {3, 32, 3, 3, 3}) && dims_are(p, "decoder.final.bias", {3}) &&
	          p.tensors.find("decoder.res_stack.2.conv1.weight") == p.tensors.end();
	for (int r = 0; r < 2 && ok; ++r) {
		const std::string pre = "decoder.res_stack." + std::to_string(r);
		for (const char* v : {".gn1.weight", ".gn1.bias", ".gn2.weight", ".gn2.bias", ".conv1.bias", ".conv2.bias"}) ok = ok && dims_are(p, pre + v, {128});
		ok = ok && dims_are(p, pre + ".conv1.weight", {128, 128, 3, 3, 3}) && dims_are(p, pre + ".conv2.weight", {128, 128, 3, 3, 3});
	}
	return ok;
}

// up_conv -> PixelShuffle3D(2) -> final, folded per output channel (decode_tc128.cuh; the same grouping of final's taps
// by (r, eps) as build_decoder_fold of the float model, decode_tc_host.cpp).
void build_decoder128_fold(const WeightPack& p, std::vector<float>& wg, std::vector<float>& bg) {
	if (!decoder128_supports(p)) throw std::runtime_error("decoder128 fold: unsupported architecture");
	const float* up_w = p.get("decoder.up_conv.weight").data;  // [256][128][27]
	const float* up_b = p.get("decoder.up_conv.bias").data;    // [256]
	const float* fin_w = p.get("decoder.final.weight").data;   // [3][32][27]
	const size_t per_n = (size_t)128 * 27;
	std::vector<double> acc((size_t)3 * 64 * per_n, 0.0), bacc(3 * 64, 0.0);
	for (int c = 0; c < 3; ++c)
		for (int r = 0; r < 8; ++r)
			for (int s2 = 0; s2 < 27; ++s2) {
				const int rr[3] = {(r >> 2) & 1, (r >> 1) & 1, r & 1};
				const int ss[3] = {s2 / 9 - 1, (s2 / 3) % 3 - 1, s2 % 3 - 1};
				int eps = 0, ru = 0;
				for (int ax = 0; ax < 3; ++ax) {
					const int t = rr[ax] + ss[ax];      // -1 .. 2
					const int e = t < 0 ? -1 : t >> 1;  // floor(t / 2)
					eps = (eps << 1) | (e != 0);
					ru = (ru << 1) | (t & 1);           // t mod 2, also for t = -1
				}
				const int n = c * 64 + r * 8 + eps;
				for (int oc = 0; oc < 32; ++oc) {
					const double f = fin_w[((size_t)c * 32 + oc) * 27 + s2];
					const int uc = oc * 8 + ru;
					bacc[n] += f * up_b[uc];
					const float* src = up_w + (size_t)uc * per_n;
					double* dst = acc.data() + (size_t)n * per_n;
					for (size_t i = 0; i < per_n; ++i) dst[i] += f * src[i];
				}
			}
	wg.resize(acc.size());
	for (size_t i = 0; i < acc.size(); ++i) wg[i] = (float)acc[i];
	bg.resize(bacc.size());
	for (size_t i = 0; i < bacc.size(); ++i) bg[i] = (float)bacc[i];
}

std::vector<uint8_t> build_decoder128_units(const WeightPack& p) {
	if (!decoder128_supports(p)) throw std::runtime_error("decoder128 units: unsupported architecture");
	std::vector<uint8_t> out((size_t)kPasses * kUnitsPerPass * kUnitBytes);
	uint8_t* u = out.data();
	for (const char* name : {"decoder.stem.0.weight", "decoder.res_stack.0.conv1.weight", "decoder.res_stack.0.conv2.weight",
	                         "decoder.res_stack.1.conv1.weight", "decoder.res_stack.1.conv2.weight"}) {
		const float* w = p.get(name).data;
		for (int h = 0; h < 2; ++h) u = fill_pass(u, w, h * 64);
	}
	std::vector<float> wg, bg;
	build_decoder128_fold(p, wg, bg);
	for (int c = 0; c < 3; ++c) u = fill_pass(u, wg.data(), c * 64);
	if (u != out.data() + out.size()) throw std::logic_error("decoder128 unit stream size mismatch");
	return out;
}

std::vector<float> build_decoder128_params(const WeightPack& p) {
	if (!decoder128_supports(p)) throw std::runtime_error("decoder128 params: unsupported architecture");
	std::vector<float> out;
	auto put = [&](const std::string& name, size_t n) {
		const PackTensor& t = p.get(name);
		if (t.numel() != n) throw std::runtime_error("decoder128 params: unexpected size of " + name);
		out.insert(out.end(), t.data, t.data + n);
	};
	put("decoder.stem.0.bias", 128);
	put("decoder.stem.1.weight", 128);
	put("decoder.stem.1.bias", 128);
	for (int r = 0; r < 2; ++r) {
		const std::string pre = "decoder.res_stack." + std::to_string(r);
		for (const char* v : {".gn1.weight", ".gn1.bias", ".conv1.bias", ".gn2.weight", ".gn2.bias", ".conv2.bias"}) put(pre + v, 128);
	}
	std::vector<float> wg, bg;
	build_decoder128_fold(p, wg, bg);
	out.insert(out.end(), bg.begin(), bg.end());  // [3][64]
	put("decoder.final.bias", 3);
	out.push_back(0.f);
	if (out.size() != 2116) throw std::logic_error("decoder128 parameter block size mismatch");
	return out;
}

}  // namespace vqvdb
