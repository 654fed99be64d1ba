// Host-side preparation of the tensor-core encoder's weight stream (see encode_tc.cuh).
#include <cmath>
#include <cstring>
#include <stdexcept>

#include "encode_tc_stream.hpp"
#include "fp16_split.hpp"

namespace vqvdb {

namespace {

// element (n, k) of one [2][N][16 B] operand block
inline void put(uint8_t* block, int N, int n, int k, uint16_t v) {
	std::memcpy(block + ((size_t)(k >> 3) * N + n) * 16 + (size_t)(k & 7) * 2, &v, 2);
}

}  // namespace

void build_encoder_vq_fold(const WeightPack& p, std::vector<float>& m, std::vector<float>& esq_fold, std::vector<float>& m_norm) {
	const PackTensor& e = p.get("quantizer.embedding");    // [256][128]
	const PackTensor& w = p.get("encoder.proj.weight");    // [128][32][1][1][1]
	const PackTensor& b = p.get("encoder.proj.bias");      // [128]
	if (e.numel() != (size_t)256 * 128 || w.numel() != (size_t)128 * 32 || b.numel() != 128)
		throw std::runtime_error("encoder VQ fold: unexpected proj / codebook shapes");
	m.assign((size_t)256 * 32, 0.f);
	esq_fold.assign(256, 0.f);
	m_norm.assign(257, 0.f);  // [256] = the maximum
	for (int k = 0; k < 256; ++k) {
		double e2 = 0.0, be = 0.0, n2 = 0.0;
		for (int d = 0; d < 128; ++d) {
			e2 += (double)e.data[k * 128 + d] * e.data[k * 128 + d];
			be += (double)e.data[k * 128 + d] * b.data[d];
		}
		for (int c = 0; c < 32; ++c) {
			double s = 0.0;
			for (int d = 0; d < 128; ++d) s += (double)e.data[k * 128 + d] * w.data[d * 32 + c];
			m[(size_t)k * 32 + c] = (float)s;
			n2 += s * s;
		}
		esq_fold[k] = (float)(e2 - 2.0 * be);
		m_norm[k] = std::nextafter((float)std::sqrt(n2), INFINITY);  // never under-estimates |M_k|
		m_norm[256] = m_norm[k] > m_norm[256] ? m_norm[k] : m_norm[256];
	}
}

std::vector<uint8_t> build_encoder_tc_units(const WeightPack& p, EncoderTcStream& tab) {
	std::vector<uint8_t> out;
	int nu = 0;
	auto begin_unit = [&](size_t bytes) -> uint8_t* {
		if (nu >= kEncTcUnits || bytes > kEncTcStageBytes || (bytes & 15)) throw std::logic_error("encoder tc unit table overflow");
		const size_t off = out.size();
		out.resize(off + bytes, 0);
		tab.off[nu] = (uint32_t)off;
		tab.bytes[nu] = (uint32_t)bytes;
		++nu;
		return out.data() + off;
	};
	auto alias_unit = [&](int src) {  // a later pass over bytes that are already in the stream
		if (nu >= kEncTcUnits) throw std::logic_error("encoder tc unit table overflow");
		tab.off[nu] = tab.off[src];
		tab.bytes[nu] = tab.bytes[src];
		++nu;
	};
	out.reserve(512 * 1024);
	// res16 conv1 / conv2: weight [16][16][3][3][3]
	for (const char* name : {"encoder.pre.3.conv1.weight", "encoder.pre.3.conv2.weight"}) {
		const float* w = p.get(name).data;
		const int first = nu;
		for (int pass = 1; pass <= kEncTcConvGroups; ++pass)  // one pass per tile group; passes 2.. re-read pass 1's bytes
		for (int kd = 0; kd < 3; ++kd) {
			if (pass > 1) {
				alias_unit(first + kd);
				continue;
			}
			uint8_t* u = begin_unit(3 * 3072);
			for (int kh = 0; kh < 3; ++kh)
				for (int part = 0; part < 2; ++part)
					for (int kw = 0; kw < 3; ++kw)
						for (int co = 0; co < 16; ++co)
							for (int ci = 0; ci < 16; ++ci)
								put(u + kh * 3072, 96, part * 48 + kw * 16 + co, ci,
								    split_part(w[((co * 16 + ci) * 27) + (kd * 3 + kh) * 3 + kw], part));
		}
	}
	// down: weight [32][16][4][4][4]; k = 2t + r per axis (t = tap of the 2x2x2 form, r = parity class of the input).
	// The two tw taps of a (td, th) pair are concatenated along N; one unit = 4 of the 8 parity classes of a pair.
	{
		const float* w = p.get("encoder.down.weight").data;
		constexpr int kPerUnit = 8 / (kEncTcDownUnits / 4);  // parity classes per unit: 4, or 2 with the four-stage ring
		for (int pair = 0; pair < 4; ++pair) {
			const int td = pair >> 1, th = pair & 1;
			for (int half = 0; half < 8 / kPerUnit; ++half) {
				uint8_t* u = begin_unit(kPerUnit * 4096);
				for (int pcl = 0; pcl < kPerUnit; ++pcl) {
					const int pc = half * kPerUnit + pcl, rd = pc >> 2, rh = (pc >> 1) & 1, rw = pc & 1;
					for (int part = 0; part < 2; ++part)
						for (int tw = 0; tw < 2; ++tw) {
							const int kd = 2 * td + rd, kh = 2 * th + rh, kw = 2 * tw + rw;
							for (int co = 0; co < 32; ++co)
								for (int ci = 0; ci < 16; ++ci)
									put(u + pcl * 4096, 128, part * 64 + tw * 32 + co, ci,
									    split_part(w[((co * 16 + ci) * 64) + (kd * 4 + kh) * 4 + kw], part));
						}
				}
			}
		}
	}
	// res32 conv1 / conv2: weight [32][32][3][3][3]
	for (const char* name : {"encoder.res_stack.0.conv1.weight", "encoder.res_stack.0.conv2.weight"}) {
		const float* w = p.get(name).data;
		for (int kk = 0; kk < 9; ++kk) {
			uint8_t* u = begin_unit(2 * 6144);
			for (int ks = 0; ks < 2; ++ks)
				for (int part = 0; part < 2; ++part)
					for (int kw = 0; kw < 3; ++kw)
						for (int co = 0; co < 32; ++co)
							for (int k = 0; k < 16; ++k)
								put(u + ks * 6144, 192, part * 96 + kw * 32 + co, k,
								    split_part(w[((co * 32 + ks * 16 + k) * 27) + kk * 3 + kw], part));
		}
	}
	// proj folded into the codebook (encode_tc_stream.hpp): M [256][32], split like the conv weights: per 16-channel k-step
	// one M_hi block and one M_lo block
	{
		std::vector<float> m, esq, mno;
		build_encoder_vq_fold(p, m, esq, mno);
		for (int ks = 0; ks < 2; ++ks) {
			uint8_t* u = kEncTcVqUnits == 2 ? begin_unit(2 * 8192) : nullptr;
			for (int part = 0; part < 2; ++part) {
				uint8_t* blk = kEncTcVqUnits == 2 ? u + part * 8192 : begin_unit(8192);  // four-stage ring: M_hi and M_lo are units of their own
				for (int code = 0; code < 256; ++code)
					for (int k = 0; k < 16; ++k) put(blk, 256, code, k, split_part(m[code * 32 + ks * 16 + k], part));
			}
		}
	}
	if (nu != kEncTcUnits) throw std::logic_error("encoder tc unit count mismatch");
	return out;
}

}  // namespace vqvdb
