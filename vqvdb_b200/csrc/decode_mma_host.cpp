// Host-side preparation of the tensor-core decoder's weight stream (see decode_mma.cuh).
#include <cstring>
#include <stdexcept>

#include "weights.hpp"

namespace vqvdb {

static uint16_t f32_to_bf16_rn(float f) {
	uint32_t u;
	std::memcpy(&u, &f, 4);
	if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);  // inf / nan pass through
	u += 0x7fffu + ((u >> 16) & 1u);
	return (uint16_t)(u >> 16);
}

// One unit = [64 n][64 k]: element (n, k) lives at byte n*128 + (((k>>3) ^ (n&7)) << 4) + (k&7)*2.
static void fill_unit(uint8_t* unit, const float* w, int cout_total, int cin_total, int oc0, int ic0, int tap) {
	(void)cout_total;
	for (int n = 0; n < 64; ++n)
		for (int k = 0; k < 64; ++k) {
			const float v = w[((size_t)(oc0 + n) * cin_total + (ic0 + k)) * 27 + tap];  // [cout][cin][kd][kh][kw]
			const uint16_t b = f32_to_bf16_rn(v);
			const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
			std::memcpy(unit + off, &b, 2);
		}
}

std::vector<uint8_t> build_decoder_units(const WeightPack& p) {
	std::vector<uint8_t> out((size_t)216 * 8192);
	uint8_t* u = out.data();
	const float* stem = p.get("decoder.stem.0.weight").data;
	for (int tap = 0; tap < 27; ++tap)
		for (int half = 0; half < 2; ++half, u += 8192) fill_unit(u, stem, 64, 128, 0, half * 64, tap);
	for (const char* name : {"decoder.res_stack.0.conv1.weight", "decoder.res_stack.0.conv2.weight"}) {
		const float* w = p.get(name).data;
		for (int tap = 0; tap < 27; ++tap, u += 8192) fill_unit(u, w, 64, 64, 0, 0, tap);
	}
	const float* up = p.get("decoder.up_conv.weight").data;
	for (int np = 0; np < 4; ++np)
		for (int tap = 0; tap < 27; ++tap, u += 8192) fill_unit(u, up, 256, 64, np * 64, 0, tap);
	if (u != out.data() + out.size()) throw std::logic_error("decoder unit stream size mismatch");
	return out;
}

std::vector<uint8_t> build_codebook_units(const WeightPack& p) {
	const PackTensor& e = p.get("quantizer.embedding");  // [256][128]
	std::vector<uint8_t> out((size_t)8 * 8192);
	uint8_t* u = out.data();
	for (int cg = 0; cg < 4; ++cg)
		for (int dh = 0; dh < 2; ++dh, u += 8192)
			for (int n = 0; n < 64; ++n)
				for (int k = 0; k < 64; ++k) {
					const uint16_t b = f32_to_bf16_rn(e.data[(size_t)(cg * 64 + n) * 128 + dh * 64 + k]);
					const size_t off = (size_t)n * 128 + ((size_t)((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
					std::memcpy(u + off, &b, 2);
				}
	return out;
}

std::vector<uint16_t> build_codebook_bf16(const WeightPack& p) {
	const PackTensor& e = p.get("quantizer.embedding");
	std::vector<uint16_t> out(e.numel());
	for (size_t i = 0; i < out.size(); ++i) out[i] = f32_to_bf16_rn(e.data[i]);
	return out;
}

}  // namespace vqvdb
