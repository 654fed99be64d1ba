// Host-side preparation of the tensor-core decoder's weight stream (see decode_tc.cuh).
#include <cstring>
#include <stdexcept>

#include "weights.hpp"

namespace vqvdb {

static constexpr int kDecUnitsWithFold = 216 + 27;  // == decode_tc.cuh (checked where the stream is uploaded)

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

// up_conv -> PixelShuffle3D(2) -> final, folded (decode_tc.cuh).  Along one axis an output voxel V = 2p + r reads the
// shuffled volume at U = V + s2, s2 in {-1, 0, 1}: U lies in cell p + e, e = floor((r + s2) / 2), at sub-position
// rU = (r + s2) mod 2.  Terms are grouped by (r, eps) with eps = (e != 0) per axis.
void build_decoder_fold(const WeightPack& p, std::vector<float>& wg /*[64][64][27]*/, std::vector<float>& bg /*[64]*/);
void build_decoder_fold(const WeightPack& p, std::vector<float>& wg, std::vector<float>& bg) {
	const PackTensor& up_w = p.get("decoder.up_conv.weight");  // [256][64][3][3][3]
	const PackTensor& up_b = p.get("decoder.up_conv.bias");    // [256]
	const PackTensor& fin_w = p.get("decoder.final.weight");   // [1][32][3][3][3]
	if (up_w.numel() != (size_t)256 * 64 * 27 || up_b.numel() != 256 || fin_w.numel() != (size_t)32 * 27)
		throw std::runtime_error("decoder fold: unexpected up_conv / final shapes");
	std::vector<double> acc((size_t)64 * 64 * 27, 0.0), bacc(64, 0.0);
	for (int r = 0; r < 8; ++r)
		for (int s2 = 0; s2 < 27; ++s2) {
			const int rr[3] = {(r >> 2) & 1, (r >> 1) & 1, r & 1};
			const int ss[3] = {s2 / 9 - 1, (s2 / 3) % 3 - 1, s2 % 3 - 1};
			int eps = 0, ru = 0;
			for (int ax = 0; ax < 3; ++ax) {
				const int t = rr[ax] + ss[ax];           // -1 .. 2
				const int e = t < 0 ? -1 : t >> 1;       // floor(t / 2)
				eps = (eps << 1) | (e != 0);
				ru = (ru << 1) | (t & 1);                // t mod 2, also for t = -1
			}
			const int n = r * 8 + eps;
			for (int oc = 0; oc < 32; ++oc) {
				const double f = fin_w.data[(size_t)oc * 27 + s2];
				const int c = oc * 8 + ru;
				bacc[n] += f * up_b.data[c];
				const float* src = up_w.data + (size_t)c * 64 * 27;
				double* dst = acc.data() + (size_t)n * 64 * 27;
				for (int i = 0; i < 64 * 27; ++i) dst[i] += f * src[i];
			}
		}
	wg.resize(acc.size());
	for (size_t i = 0; i < acc.size(); ++i) wg[i] = (float)acc[i];
	bg.resize(64);
	for (int i = 0; i < 64; ++i) bg[i] = (float)bacc[i];
}

std::vector<uint8_t> build_decoder_units(const WeightPack& p) {
	std::vector<uint8_t> out((size_t)kDecUnitsWithFold * 8192);
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
	std::vector<float> wg, bg;
	build_decoder_fold(p, wg, bg);
	for (int tap = 0; tap < 27; ++tap, u += 8192) fill_unit(u, wg.data(), 64, 64, 0, 0, tap);
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
