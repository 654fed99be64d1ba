// C-ABI of the B200 codec (include/vqvdb_b200.h): codec object, weight upload, and the chunked
// host-pointer pipeline that replaces the reference's serial per-batch
// H2D -> ~25 launches -> D2H -> memcpy sequence (TorchBackend.cpp:133-194).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vqvdb_b200.h"
#include "decode_tc.cuh"
#include "decode_tc128.cuh"
#include "encode_tc.cuh"
#include "encode_tc128.cuh"
#include "generic_model.cuh"
#include "model.cuh"
#include "weights.hpp"

namespace {

thread_local std::string g_create_error;

struct CudaError : std::runtime_error {
	cudaError_t code;
	CudaError(cudaError_t c, const char* what_call)
	    : std::runtime_error(std::string(what_call) + ": " + cudaGetErrorString(c)), code(c) {}
};
#define CUDA_TRY(call)                                  \
	do {                                                \
		cudaError_t _e = (call);                        \
		if (_e != cudaSuccess) throw CudaError(_e, #call); \
	} while (0)

constexpr int kSlots = 3;                   // pipeline depth of the host-pointer calls
#ifdef VQVDB_ENC128_GENERIC_FRONT
constexpr bool kEnc128GenericFront = true;   // bring-up aid: the 8^3 stage on the fp32 generic kernel
#else
constexpr bool kEnc128GenericFront = false;
#endif
constexpr int kEnc128LeavesPerCta = 28;       // vec3 encoder: leaves per CTA and front/back kernel pair (a batch = num_sms x 28: both grids end level)
constexpr uint32_t kDefaultChunk = 16384;   // leaves per chunk: 32 MiB of voxels, 1 MiB of indices
// Calls of at most this many leaves (the reference's SOPs hand over 64 at a time by default, 1024 / 8192 at most:
// SOP_VQVDB_Encoder.cpp:36, SOP_VQVDB_Decoder.cpp:32) skip the copy engines: the kernel reads its input from, and
// writes its result to, pinned HOST memory directly (mapped, over PCIe) — one launch and one synchronisation instead
// of copy + launch + copy + event, which is most of a 64-leaf call's time.
constexpr int64_t kZeroCopyLeaves = 2048;

// Pageable caller memory (the reference's callers hand over std::vector storage: VQVAECodec.cpp:48,114) cannot be DMA'd
// directly, so it is staged through the slots' pinned buffers.  One memcpy thread moves ~10 GB/s, a quarter of what the
// PCIe link takes and less than the decode kernel produces (2 KB x 15 M leaves/s = 31 GB/s); the staging copies are
// therefore split over a small pool of threads that lives as long as the codec.
class CopyPool {
   public:
	explicit CopyPool(int n_threads) {
		for (int i = 1; i < n_threads; ++i) workers_.emplace_back([this, i] { run(i); });
	}
	~CopyPool() {
		{
			std::lock_guard<std::mutex> lk(m_);
			stop_ = true;
			++generation_;
		}
		wake_.notify_all();
		for (auto& t : workers_) t.join();
	}
	int threads() const { return (int)workers_.size() + 1; }
	// memcpy(dst, src, bytes) by every thread of the pool (the caller is thread 0); returns when all slices are done.
	void copy(void* dst, const void* src, size_t bytes) {
		const int n = threads();
		if (n == 1 || bytes < (size_t)(1u << 20)) {
			std::memcpy(dst, src, bytes);
			return;
		}
		{
			std::lock_guard<std::mutex> lk(m_);
			dst_ = static_cast<char*>(dst);
			src_ = static_cast<const char*>(src);
			bytes_ = bytes;
			pending_ = n - 1;
			++generation_;
		}
		wake_.notify_all();
		slice(0);
		std::unique_lock<std::mutex> lk(m_);
		done_.wait(lk, [this] { return pending_ == 0; });
	}

   private:
	void slice(int i) {
		const size_t n = (size_t)threads();
		const size_t per = ((bytes_ + n - 1) / n + 4095) & ~size_t(4095);  // page-sized slices
		const size_t lo = std::min(bytes_, per * (size_t)i), hi = std::min(bytes_, lo + per);
		if (hi > lo) std::memcpy(dst_ + lo, src_ + lo, hi - lo);
	}
	void run(int i) {
		uint64_t seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(m_);
				wake_.wait(lk, [&] { return generation_ != seen; });
				seen = generation_;
				if (stop_) return;
			}
			slice(i);
			std::lock_guard<std::mutex> lk(m_);
			if (--pending_ == 0) done_.notify_one();
		}
	}
	std::vector<std::thread> workers_;
	std::mutex m_;
	std::condition_variable wake_, done_;
	uint64_t generation_ = 0;
	int pending_ = 0;
	bool stop_ = false;
	char* dst_ = nullptr;
	const char* src_ = nullptr;
	size_t bytes_ = 0;
};

int default_copy_threads() {
	if (const char* v = std::getenv("VQVDB_B200_COPY_THREADS")) return std::max(1, std::min(64, std::atoi(v)));
	const unsigned hw = std::thread::hardware_concurrency();
	return (int)std::max(1u, std::min(8u, hw / 2));
}

struct Slot {
	cudaStream_t stream = nullptr;
	cudaEvent_t done = nullptr;
	float* d_vox = nullptr;      // chunk * C * 512 floats
	uint8_t* d_idx = nullptr;    // chunk * 64 bytes
	float* h_vox = nullptr;      // pinned staging
	uint8_t* h_idx = nullptr;
	int64_t pending_first = -1;  // leaf range whose results still sit in the staging buffers
	int64_t pending_count = 0;
};

}  // namespace

struct vqvdb_b200_codec {
	int device = 0;
	int num_sms = 0;
	int channels = 1, D = 128, K = 256;
	uint32_t chunk = kDefaultChunk;
	std::string decode_path = "fp32";
	std::string last_error;
	std::atomic<uint64_t> launches{0};
	cudaStream_t compute = nullptr;
	float* arena = nullptr;  // every fp32 device weight table lives in this one allocation
	uint8_t* mma_arena = nullptr;  // bf16 weight-unit stream + bf16 codebook of the tensor-core decoder
	int decode_kind = 2;  // 1 = fp32 FFMA (checking path), 2 = bf16 tcgen05/TMEM
	int encode_kind = 2;  // 1 = fp32 FFMA, 2 = fp16x2 split on tcgen05/TMEM (fp32-level accuracy)
	std::string encode_path = "fp32";
	uint8_t* enc_tc_arena = nullptr;  // fp16 hi/lo weight-unit stream of the tensor-core encoder
	vqvdb::EncoderTcStream enc_tc{};
	vqvdb::DecoderMmaWeights dec_mma{};
	bool generic = false;            // architecture-generic kernels (vec3 model) instead of the specialised ones
	vqvdb::GenericModel gen{};
	float* gen_scratch = nullptr;
	bool dec128 = false;             // the 128-channel tensor-core decoder (decode_tc128.cu) serves this model's decode
	uint8_t* dec128_arena = nullptr;  // its bf16 unit stream, bf16 codebook and fp32 parameter block
	vqvdb::Decoder128Weights dec128_w{};
	int gen_grid = 0;
	bool enc128 = false;             // the tensor-core vec3 encoder (encode_tc128.cu) serves this model's encode
	uint8_t* enc128_arena = nullptr;  // its fp16 hi/lo unit streams, transposed codebook and fp32 parameter blocks
	float* enc128_y = nullptr;        // per pipeline slot: the stride-2 conv's output of one batch of leaves
	int64_t enc128_batch = 0;         // leaves per batch
	vqvdb::Encoder128BackWeights enc128_back{};
	vqvdb::Encoder128FrontWeights enc128_front{};
	vqvdb::Encoder128PreWeights enc128_pre{};   // pre.0 weights, passed to the front kernel by value
	float* enc128_scratch = nullptr;  // per pipeline slot: the front kernel's per-CTA fp32 scratch
	vqvdb::EncoderWeights enc{};
	vqvdb::EncoderUnits enc_units{};
	vqvdb::DecoderWeights dec{};
	Slot slots[kSlots];
	bool staging_ready = false;
	std::unique_ptr<CopyPool> copier;  // created with the staging buffers, on the first host-pointer call

	~vqvdb_b200_codec() {
		cudaSetDevice(device);
		for (auto& s : slots) {
			if (s.stream) cudaStreamSynchronize(s.stream);
			if (s.d_vox) cudaFree(s.d_vox);
			if (s.d_idx) cudaFree(s.d_idx);
			if (s.h_vox) cudaFreeHost(s.h_vox);
			if (s.h_idx) cudaFreeHost(s.h_idx);
			if (s.done) cudaEventDestroy(s.done);
			if (s.stream) cudaStreamDestroy(s.stream);
		}
		if (compute) {
			cudaStreamSynchronize(compute);
			cudaStreamDestroy(compute);
		}
		if (arena) cudaFree(arena);
		if (mma_arena) cudaFree(mma_arena);
		if (enc_tc_arena) cudaFree(enc_tc_arena);
		if (gen_scratch) cudaFree(gen_scratch);
		if (dec128_arena) cudaFree(dec128_arena);
		if (enc128_arena) cudaFree(enc128_arena);
		if (enc128_y) cudaFree(enc128_y);
		if (enc128_scratch) cudaFree(enc128_scratch);
	}
};

namespace {

using vqvdb::PackTensor;
using vqvdb::WeightPack;

// Collects host tensors, then uploads them as one arena and patches the device pointers.
struct ArenaBuilder {
	std::vector<float> host;
	std::vector<std::pair<const float**, size_t>> fixups;
	void add(const float** slot, const float* data, size_t n) {
		const size_t off = (host.size() + 63) & ~size_t(63);  // 256-byte aligned tables
		host.resize(off + n);
		std::memcpy(host.data() + off, data, n * sizeof(float));
		fixups.emplace_back(slot, off);
	}
	void add(const float** slot, const std::vector<float>& v) { add(slot, v.data(), v.size()); }
	void add(const float** slot, const PackTensor& t) { add(slot, t.data, t.numel()); }
	float* upload() {
		float* dev = nullptr;
		CUDA_TRY(cudaMalloc(&dev, host.size() * sizeof(float)));
		cudaError_t e = cudaMemcpy(dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
		if (e != cudaSuccess) {
			cudaFree(dev);
			throw CudaError(e, "cudaMemcpy(weights)");
		}
		for (auto& f : fixups) *f.first = dev + f.second;
		return dev;
	}
};

void expect_dims(const WeightPack& p, const std::string& name, std::initializer_list<int> dims) {
	const PackTensor& t = p.get(name);
	if (t.dims != std::vector<int>(dims)) throw std::runtime_error("weight pack: unexpected shape for " + name);
}

void add_res(ArenaBuilder& ab, const WeightPack& p, const std::string& prefix, vqvdb::ResWeights& r) {
	ab.add(&r.gn1_w, p.get(prefix + ".gn1.weight"));
	ab.add(&r.gn1_b, p.get(prefix + ".gn1.bias"));
	ab.add(&r.c1_w, vqvdb::transpose_conv_weight(p.get(prefix + ".conv1.weight")));
	ab.add(&r.c1_b, p.get(prefix + ".conv1.bias"));
	ab.add(&r.gn2_w, p.get(prefix + ".gn2.weight"));
	ab.add(&r.gn2_b, p.get(prefix + ".gn2.bias"));
	ab.add(&r.c2_w, vqvdb::transpose_conv_weight(p.get(prefix + ".conv2.weight")));
	ab.add(&r.c2_b, p.get(prefix + ".conv2.bias"));
}

void upload_float_model(vqvdb_b200_codec& c, const WeightPack& p) {
	// Architecture check: the kernels are specialised to the shipped float model.
	expect_dims(p, "encoder.pre.0.weight", {16, 1, 3, 3, 3});
	expect_dims(p, "encoder.pre.3.conv1.weight", {16, 16, 3, 3, 3});
	expect_dims(p, "encoder.down.weight", {32, 16, 4, 4, 4});
	expect_dims(p, "encoder.res_stack.0.conv1.weight", {32, 32, 3, 3, 3});
	expect_dims(p, "encoder.attn.fc.0.weight", {8, 32});
	expect_dims(p, "encoder.attn.fc.2.weight", {32, 8});
	expect_dims(p, "encoder.proj.weight", {128, 32, 1, 1, 1});
	expect_dims(p, "quantizer.embedding", {256, 128});
	expect_dims(p, "decoder.stem.0.weight", {64, 128, 3, 3, 3});
	expect_dims(p, "decoder.res_stack.0.conv1.weight", {64, 64, 3, 3, 3});
	expect_dims(p, "decoder.attn.fc.0.weight", {16, 64});
	expect_dims(p, "decoder.attn.fc.2.weight", {64, 16});
	expect_dims(p, "decoder.up_conv.weight", {256, 64, 3, 3, 3});
	expect_dims(p, "decoder.final.weight", {1, 32, 3, 3, 3});

	ArenaBuilder ab;
	auto& e = c.enc;
	ab.add(&e.pre_w, vqvdb::transpose_conv_weight(p.get("encoder.pre.0.weight")));
	ab.add(&e.pre_b, p.get("encoder.pre.0.bias"));
	ab.add(&e.pre_gn_w, p.get("encoder.pre.1.weight"));
	ab.add(&e.pre_gn_b, p.get("encoder.pre.1.bias"));
	add_res(ab, p, "encoder.pre.3", e.res16);
	ab.add(&e.down_w, vqvdb::transpose_conv_weight(p.get("encoder.down.weight")));
	ab.add(&e.down_b, p.get("encoder.down.bias"));
	add_res(ab, p, "encoder.res_stack.0", e.res32);
	ab.add(&e.fc0, p.get("encoder.attn.fc.0.weight"));
	ab.add(&e.fc2, p.get("encoder.attn.fc.2.weight"));
	ab.add(&e.proj_w, vqvdb::transpose_conv_weight(p.get("encoder.proj.weight")));
	ab.add(&e.proj_b, p.get("encoder.proj.bias"));
	const PackTensor& emb = p.get("quantizer.embedding");
	std::vector<float> emb_t((size_t)128 * 256), emb_sq(256);
	for (int k = 0; k < 256; ++k) {
		float s = 0.f;  // torch.sum(embedding ** 2, dim=1), save_for_inference.py:58
		for (int d = 0; d < 128; ++d) {
			const float v = emb.data[k * 128 + d];
			emb_t[(size_t)d * 256 + k] = v;
			s += v * v;
		}
		emb_sq[k] = s;
	}
	std::vector<float> emb_norm(256);
	for (int k = 0; k < 256; ++k) {
		double s2 = 0.0;
		for (int d = 0; d < 128; ++d) s2 += (double)emb.data[k * 128 + d] * emb.data[k * 128 + d];
		emb_norm[k] = std::nextafter((float)std::sqrt(s2), INFINITY);  // never under-estimates |e_k|
	}
	ab.add(&e.emb_t, emb_t);
	ab.add(&e.emb_sq, emb_sq);
	ab.add(&e.emb_norm, emb_norm);
	{
		std::vector<float> fm, fesq, fnorm;
		vqvdb::build_encoder_vq_fold(p, fm, fesq, fnorm);
		ab.add(&e.fold_esq, fesq);
		ab.add(&e.fold_norm, fnorm);
	}
	ab.add(&e.emb, emb);

	auto& d = c.dec;
	ab.add(&d.emb, emb);
	ab.add(&d.stem_w, vqvdb::transpose_conv_weight(p.get("decoder.stem.0.weight")));
	ab.add(&d.stem_b, p.get("decoder.stem.0.bias"));
	ab.add(&d.stem_gn_w, p.get("decoder.stem.1.weight"));
	ab.add(&d.stem_gn_b, p.get("decoder.stem.1.bias"));
	add_res(ab, p, "decoder.res_stack.0", d.res64);
	ab.add(&d.fc0, p.get("decoder.attn.fc.0.weight"));
	ab.add(&d.fc2, p.get("decoder.attn.fc.2.weight"));
	ab.add(&d.up_w, vqvdb::transpose_conv_weight(p.get("decoder.up_conv.weight")));
	ab.add(&d.up_b, p.get("decoder.up_conv.bias"));
	ab.add(&d.fin_w, vqvdb::transpose_conv_weight(p.get("decoder.final.weight")));
	ab.add(&d.fin_b, p.get("decoder.final.bias"));
	{
		std::vector<float> wg, bg;
		vqvdb::build_decoder_fold(p, wg, bg);
		ab.add(&c.dec_mma.fold_b, bg);
	}
	c.arena = ab.upload();

	// tensor-core decoder: bf16 unit stream + bf16 codebook; fp32 vectors are shared with the fp32 path.
	// The encoder's VQ shortlist pass reads the codebook as 8 more bf16 units from the same allocation.
	const std::vector<uint8_t> units = vqvdb::build_decoder_units(p);
	if (units.size() != (size_t)vqvdb::kDecUnitsWithFold * 8192) throw std::logic_error("decoder unit stream size mismatch");
	const std::vector<uint16_t> cb = vqvdb::build_codebook_bf16(p);
	const std::vector<uint8_t> cb_units = vqvdb::build_codebook_units(p);
	CUDA_TRY(cudaMalloc(&c.mma_arena, units.size() + cb.size() * 2 + cb_units.size()));
	CUDA_TRY(cudaMemcpy(c.mma_arena, units.data(), units.size(), cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(c.mma_arena + units.size(), cb.data(), cb.size() * 2, cudaMemcpyHostToDevice));
	CUDA_TRY(cudaMemcpy(c.mma_arena + units.size() + cb.size() * 2, cb_units.data(), cb_units.size(), cudaMemcpyHostToDevice));

	// encoder weight-unit stream: slices of the transposed fp32 tables, then the bf16 codebook tiles,
	// in the order the kernel consumes them
	{
		auto& U = c.enc_units;
		int n = 0;
		auto add = [&](const void* ptr, size_t bytes) {
			if (n >= vqvdb::kEncUnits) throw std::logic_error("encoder unit table overflow");
			U.ptr[n] = ptr;
			U.bytes[n] = (uint32_t)bytes;
			if (bytes > 8192 || (bytes & 15) || (reinterpret_cast<uintptr_t>(ptr) & 15)) throw std::logic_error("encoder unit misaligned");
			++n;
		};
		add(e.pre_w, 27 * 16 * 4);
		for (const float* t : {e.res16.c1_w, e.res16.c2_w})
			for (int q = 0; q < 4; ++q) add(t + (size_t)q * 4 * 27 * 16, 4 * 27 * 16 * 4);
		add(e.pre_w, 27 * 16 * 4);  // the 8^3 residual is re-derived from the input leaf after conv2
		for (int ic = 0; ic < 16; ++ic) add(e.down_w + (size_t)ic * 64 * 32, 64 * 32 * 4);
		for (const float* t : {e.res32.c1_w, e.res32.c2_w})
			for (int q = 0; q < 16; ++q) add(t + (size_t)q * 2 * 27 * 32, 2 * 27 * 32 * 4);
		for (int q = 0; q < 2; ++q) add(e.proj_w + (size_t)q * 16 * 128, 16 * 128 * 4);
		const uint8_t* cbu = c.mma_arena + units.size() + cb.size() * 2;
		for (int round = 0; round < 2; ++round)
			for (int q = 0; q < 8; ++q) add(cbu + (size_t)q * 8192, 8192);
		if (n != vqvdb::kEncUnits) throw std::logic_error("encoder unit table size mismatch");
	}
	// tensor-core encoder: fp16 hi/lo unit stream (encode_tc.cuh)
	{
		const std::vector<uint8_t> eu = vqvdb::build_encoder_tc_units(p, c.enc_tc);
		CUDA_TRY(cudaMalloc(&c.enc_tc_arena, eu.size()));
		CUDA_TRY(cudaMemcpy(c.enc_tc_arena, eu.data(), eu.size(), cudaMemcpyHostToDevice));
		c.enc_tc.units = c.enc_tc_arena;
	}
	auto& m = c.dec_mma;
	m.units = c.mma_arena;
	m.emb_bf16 = reinterpret_cast<const __nv_bfloat16*>(c.mma_arena + units.size());
	m.stem_b = d.stem_b;
	m.stem_gn_w = d.stem_gn_w;
	m.stem_gn_b = d.stem_gn_b;
	m.res = d.res64;
	m.fc0 = d.fc0;
	m.fc2 = d.fc2;
	m.up_b = d.up_b;
	m.fin_w = d.fin_w;
	m.fin_b = d.fin_b;
}

void add_gres(ArenaBuilder& ab, const WeightPack& p, const std::string& prefix, vqvdb::GenericRes& r) {
	ab.add(&r.gn1_w, p.get(prefix + ".gn1.weight"));
	ab.add(&r.gn1_b, p.get(prefix + ".gn1.bias"));
	ab.add(&r.c1_w, vqvdb::transpose_conv_weight(p.get(prefix + ".conv1.weight")));
	ab.add(&r.c1_b, p.get(prefix + ".conv1.bias"));
	ab.add(&r.gn2_w, p.get(prefix + ".gn2.weight"));
	ab.add(&r.gn2_b, p.get(prefix + ".gn2.bias"));
	ab.add(&r.c2_w, vqvdb::transpose_conv_weight(p.get(prefix + ".conv2.weight")));
	ab.add(&r.c2_b, p.get(prefix + ".conv2.bias"));
}

// The reference's vec3 architecture (EncoderVec3 / DecoderVec3, python/VQVAE_v2.py:278-325), described at run time.
void upload_generic_model(vqvdb_b200_codec& c, const WeightPack& p) {
	auto& m = c.gen;
	m.cin = p.in_channels;
	m.D = p.embedding_dim;
	m.K = p.num_embeddings;
	const bool vec3 = m.cin != 1;
	const std::string down = vec3 ? "encoder.down1" : "encoder.down";
	const PackTensor& dw = p.get(down + ".weight");
	auto rank_of = [&](const char* name, size_t rank) {
		if (p.get(name).dims.size() != rank) throw std::runtime_error(std::string("weight pack: unexpected rank for ") + name);
	};
	if (dw.dims.size() != 5) throw std::runtime_error("weight pack: unexpected rank for " + down + ".weight");
	rank_of("encoder.attn.fc.0.weight", 2);
	rank_of("decoder.stem.0.weight", 5);
	rank_of("decoder.attn.fc.0.weight", 2);
	m.e_c0 = dw.dims[1];
	m.e_c1 = dw.dims[0];
	m.e_down_k = dw.dims[2];
	m.e_gn0 = vec3 ? 8 : 4;
	m.e_nres = vec3 ? 2 : 1;
	m.e_red = p.get("encoder.attn.fc.0.weight").dims[0];
	m.d_c = p.get("decoder.stem.0.weight").dims[0];
	m.d_nres = vec3 ? 2 : 1;
	m.d_red = p.get("decoder.attn.fc.0.weight").dims[0];
	if (m.e_c0 > 64 || m.e_c1 > 128 || m.D > 128 || m.d_c > 128 || m.e_red > 64 || m.d_red > 64)
		throw std::runtime_error("weight pack: architecture outside the generic kernels' limits");
	// uint8 indices address the whole codebook: with K < 256 an index from an untrusted .vqvdb file could read past it
	// (F.embedding in the reference raises on such an index; save_for_inference.py:63-64)
	if (m.K != 256) throw std::runtime_error("weight pack: the generic kernels need a 256-entry codebook (uint8 indices)");
	// every tensor's shape against the architecture the kernels index by (python/VQVAE_v2.py:278-325)
	const int e0 = m.e_c0, e1 = m.e_c1, dc = m.d_c, dk = m.e_down_k;
	expect_dims(p, "encoder.pre.0.weight", {e0, m.cin, 3, 3, 3});
	expect_dims(p, "encoder.pre.0.bias", {e0});
	expect_dims(p, "encoder.pre.1.weight", {e0});
	expect_dims(p, "encoder.pre.1.bias", {e0});
	auto expect_res = [&](const std::string& prefix, int ch) {
		for (const char* gn : {".gn1", ".gn2"}) {
			expect_dims(p, prefix + gn + ".weight", {ch});
			expect_dims(p, prefix + gn + ".bias", {ch});
		}
		for (const char* cv : {".conv1", ".conv2"}) {
			expect_dims(p, prefix + cv + ".weight", {ch, ch, 3, 3, 3});
			expect_dims(p, prefix + cv + ".bias", {ch});
		}
	};
	expect_res("encoder.pre.3", e0);
	expect_dims(p, down + ".weight", {e1, e0, dk, dk, dk});
	expect_dims(p, down + ".bias", {e1});
	if (dk != 3 && dk != 4) throw std::runtime_error("weight pack: unsupported kernel size of the stride-2 conv");
	for (int r = 0; r < m.e_nres; ++r) expect_res("encoder.res_stack." + std::to_string(r), e1);
	expect_dims(p, "encoder.attn.fc.0.weight", {m.e_red, e1});
	expect_dims(p, "encoder.attn.fc.2.weight", {e1, m.e_red});
	expect_dims(p, "encoder.proj.weight", {m.D, e1, 1, 1, 1});
	expect_dims(p, "encoder.proj.bias", {m.D});
	expect_dims(p, "quantizer.embedding", {m.K, m.D});
	expect_dims(p, "decoder.stem.0.weight", {dc, m.D, 3, 3, 3});
	expect_dims(p, "decoder.stem.0.bias", {dc});
	expect_dims(p, "decoder.stem.1.weight", {dc});
	expect_dims(p, "decoder.stem.1.bias", {dc});
	for (int r = 0; r < m.d_nres; ++r) expect_res("decoder.res_stack." + std::to_string(r), dc);
	expect_dims(p, "decoder.attn.fc.0.weight", {m.d_red, dc});
	expect_dims(p, "decoder.attn.fc.2.weight", {dc, m.d_red});
	expect_dims(p, "decoder.up_conv.weight", {256, dc, 3, 3, 3});
	expect_dims(p, "decoder.up_conv.bias", {256});
	expect_dims(p, "decoder.final.weight", {m.cin, 32, 3, 3, 3});
	expect_dims(p, "decoder.final.bias", {m.cin});
	ArenaBuilder ab;
	ab.add(&m.e_pre_w, vqvdb::transpose_conv_weight(p.get("encoder.pre.0.weight")));
	ab.add(&m.e_pre_b, p.get("encoder.pre.0.bias"));
	ab.add(&m.e_gn_w, p.get("encoder.pre.1.weight"));
	ab.add(&m.e_gn_b, p.get("encoder.pre.1.bias"));
	add_gres(ab, p, "encoder.pre.3", m.e_res0);
	ab.add(&m.e_down_w, vqvdb::transpose_conv_weight(dw));
	ab.add(&m.e_down_b, p.get(down + ".bias"));
	for (int r = 0; r < m.e_nres; ++r) add_gres(ab, p, "encoder.res_stack." + std::to_string(r), m.e_res[r]);
	ab.add(&m.e_fc0, p.get("encoder.attn.fc.0.weight"));
	ab.add(&m.e_fc2, p.get("encoder.attn.fc.2.weight"));
	auto transposed2d = [](const PackTensor& t) {  // [rows][cols] -> [cols][rows]
		const int rows = t.dims[0], cols = t.dims[1];
		std::vector<float> out((size_t)rows * cols);
		for (int r = 0; r < rows; ++r)
			for (int c2 = 0; c2 < cols; ++c2) out[(size_t)c2 * rows + r] = t.data[(size_t)r * cols + c2];
		return out;
	};
	ab.add(&m.e_fc2_t, transposed2d(p.get("encoder.attn.fc.2.weight")));
	ab.add(&m.d_fc2_t, transposed2d(p.get("decoder.attn.fc.2.weight")));
	ab.add(&m.e_proj_w, vqvdb::transpose_conv_weight(p.get("encoder.proj.weight")));
	ab.add(&m.e_proj_b, p.get("encoder.proj.bias"));
	const PackTensor& emb = p.get("quantizer.embedding");
	std::vector<float> emb_sq((size_t)m.K);
	for (int k = 0; k < m.K; ++k) {
		float s = 0.f;
		for (int d = 0; d < m.D; ++d) s += emb.data[(size_t)k * m.D + d] * emb.data[(size_t)k * m.D + d];
		emb_sq[k] = s;
	}
	ab.add(&m.emb, emb);
	ab.add(&m.emb_sq, emb_sq);
	ab.add(&m.d_stem_w, vqvdb::transpose_conv_weight(p.get("decoder.stem.0.weight")));
	ab.add(&m.d_stem_b, p.get("decoder.stem.0.bias"));
	ab.add(&m.d_gn_w, p.get("decoder.stem.1.weight"));
	ab.add(&m.d_gn_b, p.get("decoder.stem.1.bias"));
	for (int r = 0; r < m.d_nres; ++r) add_gres(ab, p, "decoder.res_stack." + std::to_string(r), m.d_res[r]);
	ab.add(&m.d_fc0, p.get("decoder.attn.fc.0.weight"));
	ab.add(&m.d_fc2, p.get("decoder.attn.fc.2.weight"));
	ab.add(&m.d_up_w, vqvdb::transpose_conv_weight(p.get("decoder.up_conv.weight")));
	ab.add(&m.d_up_b, p.get("decoder.up_conv.bias"));
	ab.add(&m.d_fin_w, vqvdb::transpose_conv_weight(p.get("decoder.final.weight")));
	ab.add(&m.d_fin_b, p.get("decoder.final.bias"));
	c.arena = ab.upload();
	c.gen_grid = 3 * c.num_sms;  // 80 registers x 256 threads: three CTAs per SM
	// one scratch area per pipeline slot plus one for the device-pointer entry points (slot index kSlots)
	CUDA_TRY(cudaMalloc(&c.gen_scratch, (kSlots + 1) * vqvdb::generic_scratch_floats(c.gen_grid) * sizeof(float)));
	c.generic = true;
	// the reference's vec3 decoder (128 channels, two residual blocks) has a tensor-core path of its own
	if (vqvdb::decoder128_supports(p)) {
		const std::vector<uint8_t> units = vqvdb::build_decoder128_units(p);
		const std::vector<uint16_t> cb = vqvdb::build_codebook_bf16(p);
		const std::vector<float> par = vqvdb::build_decoder128_params(p);
		const size_t off_cb = units.size(), off_par = (off_cb + cb.size() * 2 + 255) & ~size_t(255);
		CUDA_TRY(cudaMalloc(&c.dec128_arena, off_par + par.size() * sizeof(float)));
		CUDA_TRY(cudaMemcpy(c.dec128_arena, units.data(), units.size(), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(c.dec128_arena + off_cb, cb.data(), cb.size() * 2, cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(c.dec128_arena + off_par, par.data(), par.size() * sizeof(float), cudaMemcpyHostToDevice));
		c.dec128_w.units = c.dec128_arena;
		c.dec128_w.emb_bf16 = reinterpret_cast<const __nv_bfloat16*>(c.dec128_arena + off_cb);
		c.dec128_w.par = reinterpret_cast<const float*>(c.dec128_arena + off_par);
		c.dec128_w.fc0 = m.d_fc0;
		c.dec128_w.fc2_t = m.d_fc2_t;
		c.dec128 = true;
	}
	// ... and so has its encoder (64 / 128 channels, one + two residual blocks)
	if (vqvdb::encoder128_supports(p)) {
		const std::vector<uint8_t> back = vqvdb::build_encoder128_back_units(p);
		const std::vector<float> back_par = vqvdb::build_encoder128_back_params(p);
		const std::vector<uint8_t> vq_units = vqvdb::build_encoder128_vq_units(p);  // proj folded into the codebook, fp16 hi / lo
		const std::vector<uint8_t> front = vqvdb::build_encoder128_front_units(p);
		const std::vector<float> front_par = vqvdb::build_encoder128_front_params(p);
		const size_t off_par = back.size(), off_emb = (off_par + back_par.size() * sizeof(float) + 255) & ~size_t(255);
		const size_t off_front = (off_emb + vq_units.size() + 1023) & ~size_t(1023), off_fpar = off_front + front.size();
		CUDA_TRY(cudaMalloc(&c.enc128_arena, off_fpar + front_par.size() * sizeof(float)));
		CUDA_TRY(cudaMemcpy(c.enc128_arena + off_front, front.data(), front.size(), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(c.enc128_arena + off_fpar, front_par.data(), front_par.size() * sizeof(float), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMalloc(&c.enc128_scratch, (size_t)(kSlots + 1) * vqvdb::encode_tc128_front_scratch_floats(c.num_sms) * sizeof(float)));
		c.enc128_front.units = c.enc128_arena + off_front;
		c.enc128_front.par = reinterpret_cast<const float*>(c.enc128_arena + off_fpar);
		{
			const std::vector<float> pw = vqvdb::transpose_conv_weight(p.get("encoder.pre.0.weight"));  // [3 ic][27][64 c]
			if (pw.size() != sizeof(c.enc128_pre.w) / sizeof(float)) throw std::logic_error("encoder128: pre.0 weight size mismatch");
			std::memcpy(c.enc128_pre.w, pw.data(), sizeof(c.enc128_pre.w));
		}
		CUDA_TRY(cudaMemcpy(c.enc128_arena, back.data(), back.size(), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(c.enc128_arena + off_par, back_par.data(), back_par.size() * sizeof(float), cudaMemcpyHostToDevice));
		CUDA_TRY(cudaMemcpy(c.enc128_arena + off_emb, vq_units.data(), vq_units.size(), cudaMemcpyHostToDevice));
		c.enc128_batch = (int64_t)c.num_sms * kEnc128LeavesPerCta;
		CUDA_TRY(cudaMalloc(&c.enc128_y, (size_t)(kSlots + 1) * c.enc128_batch * 8192 * sizeof(float)));
		c.enc128_back.units = c.enc128_arena;
		c.enc128_back.par = reinterpret_cast<const float*>(c.enc128_arena + off_par);
		c.enc128_back.fc0 = m.e_fc0;
		c.enc128_back.fc2_t = m.e_fc2_t;
		c.enc128_back.vq_units = c.enc128_arena + off_emb;
		c.enc128_back.proj_t = m.e_proj_w;
		c.enc128_back.emb = m.emb;
		c.enc128_back.emb_sq = m.emb_sq;
		c.enc128 = true;
	}
}

void ensure_staging(vqvdb_b200_codec& c) {
	if (c.staging_ready) return;
	const size_t vox_bytes = (size_t)c.chunk * c.channels * 512 * sizeof(float);
	const size_t idx_bytes = (size_t)c.chunk * 64;
	for (auto& s : c.slots) {  // a failed attempt (out of memory part-way) is resumed, not repeated over live handles
		if (!s.stream) CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
		if (!s.done) CUDA_TRY(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
		if (!s.d_vox) CUDA_TRY(cudaMalloc(&s.d_vox, vox_bytes));
		if (!s.d_idx) CUDA_TRY(cudaMalloc(&s.d_idx, idx_bytes));
		if (!s.h_vox) CUDA_TRY(cudaMallocHost(&s.h_vox, vox_bytes));
		if (!s.h_idx) CUDA_TRY(cudaMallocHost(&s.h_idx, idx_bytes));
	}
	if (!c.copier) c.copier = std::make_unique<CopyPool>(default_copy_threads());
	c.staging_ready = true;
}

bool is_pinned_host(const void* p) {
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return a.type == cudaMemoryTypeHost;
}

// Device-side alias of pinned host memory (cudaMallocHost / cudaHostRegister'd / torch pin_memory), or null for
// pageable memory.  One attribute query, which is not an error for an unregistered pointer.
template <class T>
T* device_alias(T* host_ptr, size_t align) {
	if (reinterpret_cast<uintptr_t>(host_ptr) & (align - 1)) return nullptr;
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
	return static_cast<T*>(a.devicePointer);
}

int fail(vqvdb_b200_codec* c, int code, const std::string& msg) {
	if (c) c->last_error = msg;
	else g_create_error = msg;
	return code;
}

int translate(vqvdb_b200_codec* c, const std::exception& e) {
	if (auto* ce = dynamic_cast<const CudaError*>(&e))
		return fail(c, ce->code == cudaErrorMemoryAllocation ? VQVDB_B200_ERR_OUT_OF_MEMORY : VQVDB_B200_ERR_CUDA, e.what());
	return fail(c, VQVDB_B200_ERR_CUDA, e.what());
}

void launch_encode(vqvdb_b200_codec& c, const float* d_leaves, int64_t n, uint8_t* d_idx, cudaStream_t st, int slot = kSlots) {
	if (c.generic && c.enc128 && c.encode_kind == 2) {
		// two kernels per batch of leaves, the stride-2 conv's output (32 KB per leaf) handed over in global memory
		float* scratch = c.enc128_scratch + (size_t)slot * vqvdb::encode_tc128_front_scratch_floats(c.num_sms);
		float* y = c.enc128_y + (size_t)slot * c.enc128_batch * 8192;
		for (int64_t first = 0; first < n; first += c.enc128_batch) {
			const int64_t nb = std::min<int64_t>(c.enc128_batch, n - first);
			if (kEnc128GenericFront) {
				float* gscratch = c.gen_scratch + (size_t)slot * vqvdb::generic_scratch_floats(c.gen_grid);
				CUDA_TRY(vqvdb::launch_encode_generic_front(c.gen, d_leaves + first * 1536, nb, y, gscratch, c.gen_grid, st));
			} else CUDA_TRY(vqvdb::launch_encode_tc128_front(c.enc128_front, c.enc128_pre, d_leaves + first * 1536, nb, y, scratch, c.num_sms, st));
			CUDA_TRY(vqvdb::launch_encode_tc128_back(c.enc128_back, y, nb, d_idx + first * 64, c.num_sms, st));
			c.launches.fetch_add(2, std::memory_order_relaxed);
		}
		return;
	}
	if (c.generic) {
		float* scratch = c.gen_scratch + (size_t)slot * vqvdb::generic_scratch_floats(c.gen_grid);
		CUDA_TRY(vqvdb::launch_encode_generic(c.gen, d_leaves, n, d_idx, scratch, c.gen_grid, st));
		if (n > 0) c.launches.fetch_add(1, std::memory_order_relaxed);
		return;
	}
	if (c.encode_kind == 2) CUDA_TRY(vqvdb::launch_encode_tc(c.enc, c.enc_tc, d_leaves, n, d_idx, c.num_sms, st));
	else CUDA_TRY(vqvdb::launch_encode_fp32(c.enc, c.enc_units, d_leaves, n, d_idx, c.num_sms, st));
	if (n > 0) c.launches.fetch_add(1, std::memory_order_relaxed);
}

void launch_decode(vqvdb_b200_codec& c, const uint8_t* d_idx, int64_t n, float* d_vox, cudaStream_t st, int slot = kSlots) {
	if (c.generic && c.dec128 && c.decode_kind == 2) {
		CUDA_TRY(vqvdb::launch_decode_tc128(c.dec128_w, d_idx, n, d_vox, c.num_sms, st));
		if (n > 0) c.launches.fetch_add(1, std::memory_order_relaxed);
		return;
	}
	if (c.generic) {
		float* scratch = c.gen_scratch + (size_t)slot * vqvdb::generic_scratch_floats(c.gen_grid);
		CUDA_TRY(vqvdb::launch_decode_generic(c.gen, d_idx, n, d_vox, scratch, c.gen_grid, st));
		if (n > 0) c.launches.fetch_add(1, std::memory_order_relaxed);
		return;
	}
	if (c.decode_kind == 2) CUDA_TRY(vqvdb::launch_decode_tc(c.dec_mma, d_idx, n, d_vox, c.num_sms, st));
	else CUDA_TRY(vqvdb::launch_decode_fp32(c.dec, d_idx, n, d_vox, c.num_sms, st));
	if (n > 0) c.launches.fetch_add(1, std::memory_order_relaxed);
}

// Drains a slot: waits for its chunk and, if results were staged, copies them to the caller's buffer.
template <class T>
void retire(CopyPool& cp, Slot& s, T* user_out, size_t elems_per_leaf, const T* staged, bool direct) {
	if (s.pending_first < 0) return;
	CUDA_TRY(cudaEventSynchronize(s.done));
	if (!direct)
		cp.copy(user_out + (size_t)s.pending_first * elems_per_leaf, staged,
		        (size_t)s.pending_count * elems_per_leaf * sizeof(T));
	s.pending_first = -1;
	s.pending_count = 0;
}

}  // namespace

extern "C" {

const char* vqvdb_b200_version(void) { return "vqvdb_b200 0.1 (sm_100a)"; }

int vqvdb_b200_create(const vqvdb_b200_config* cfg, vqvdb_b200_codec** out) {
	if (!out) return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "out is NULL");
	*out = nullptr;
	vqvdb_b200_config conf{};
	if (cfg) {
		if (cfg->struct_size < 32 || cfg->struct_size > sizeof(vqvdb_b200_config))
			return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "vqvdb_b200_config.struct_size is not set");
		std::memcpy(&conf, cfg, cfg->struct_size);
	}
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
		cudaGetLastError();
		return fail(nullptr, VQVDB_B200_ERR_NO_DEVICE, "no CUDA device is visible; this backend has no CPU fallback");
	}
	if (conf.device < 0 || conf.device >= n_dev)
		return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "device ordinal out of range");
	std::unique_ptr<vqvdb_b200_codec> c(new vqvdb_b200_codec());
	try {
		c->device = conf.device;
		CUDA_TRY(cudaSetDevice(c->device));
		cudaDeviceProp prop{};
		CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
		if (prop.major != 10)
			return fail(nullptr, VQVDB_B200_ERR_NO_DEVICE,
			            std::string("device '") + prop.name + "' is not an sm_100 (Blackwell B200) part; kernels are built for sm_100a only");
		c->num_sms = prop.multiProcessorCount;
		c->chunk = conf.chunk_leaves ? conf.chunk_leaves : kDefaultChunk;

		WeightPack pack;
		try {
			if (conf.weights_data && conf.weights_size) pack.parse(conf.weights_data, (size_t)conf.weights_size);
			else if (conf.onnx_encoder_path && conf.onnx_decoder_path) {
				const auto blob = vqvdb::onnx_to_pack(vqvdb::read_file_bytes(conf.onnx_encoder_path), vqvdb::read_file_bytes(conf.onnx_decoder_path));
				pack.parse(blob.data(), blob.size());
			} else if (conf.weights_path && conf.weights_path[0] && std::filesystem::is_directory(conf.weights_path)) {
				const std::filesystem::path dir(conf.weights_path);  // OnnxBackendFactory.cpp:97-119: <dir>/encoder.onnx + <dir>/decoder.onnx
				const auto blob = vqvdb::onnx_to_pack(vqvdb::read_file_bytes((dir / "encoder.onnx").string()), vqvdb::read_file_bytes((dir / "decoder.onnx").string()));
				pack.parse(blob.data(), blob.size());
			} else if (conf.weights_path && conf.weights_path[0]) pack.load_file(conf.weights_path);
			else pack.parse(vqvdb::vqvdb_b200_embedded_pack,
				            (size_t)(vqvdb::vqvdb_b200_embedded_pack_end - vqvdb::vqvdb_b200_embedded_pack));
			c->channels = pack.in_channels;
			c->D = pack.embedding_dim;
			c->K = pack.num_embeddings;
			if (c->channels == 1 && c->D == 128 && c->K == 256) upload_float_model(*c, pack);
			else if (c->channels == 3) upload_generic_model(*c, pack);  // vec3 architecture: generic fp32 kernels
			else return fail(nullptr, VQVDB_B200_ERR_UNSUPPORTED, "unsupported model: expected the float (C=1, D=128, K=256) or vec3 (C=3) architecture");
		} catch (const CudaError&) {
			throw;
		} catch (const std::exception& e) {
			return fail(nullptr, VQVDB_B200_ERR_BAD_WEIGHTS, e.what());
		}
		if (conf.decode_precision > VQVDB_B200_DECODE_BF16_TC)
			return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "unknown decode_precision");
		c->decode_kind = conf.decode_precision == VQVDB_B200_DECODE_DEFAULT ? (int)VQVDB_B200_DECODE_DEFAULT_KIND : (int)conf.decode_precision;
		if (conf.encode_precision > VQVDB_B200_ENCODE_FP16X2_TC)
			return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "unknown encode_precision");
		c->encode_kind = conf.encode_precision == VQVDB_B200_ENCODE_DEFAULT ? (int)VQVDB_B200_ENCODE_DEFAULT_KIND : (int)conf.encode_precision;
		c->encode_path = c->generic ? (c->enc128 && c->encode_kind == 2 ? "fp16x2_tcgen05_c128" : "fp32_generic") : c->encode_kind == 2 ? "fp16x2_tcgen05" : "fp32";
		c->decode_path = c->generic ? (c->dec128 && c->decode_kind == 2 ? "bf16_tcgen05_c128_fold" : "fp32_generic") : c->decode_kind == 2 ? "bf16_tcgen05_n192_fold" : "fp32";
		CUDA_TRY(cudaStreamCreateWithFlags(&c->compute, cudaStreamNonBlocking));
		CUDA_TRY(vqvdb::configure_encode_fp32());
		CUDA_TRY(vqvdb::configure_encode_tc());
		CUDA_TRY(vqvdb::configure_decode_fp32());
		CUDA_TRY(vqvdb::configure_decode_tc());
		CUDA_TRY(vqvdb::configure_decode_tc128());
		CUDA_TRY(vqvdb::configure_encode_tc128());
		CUDA_TRY(vqvdb::configure_encode_tc128_front());
	} catch (const std::exception& e) {
		return translate(nullptr, e);
	}
	*out = c.release();
	return VQVDB_B200_OK;
}

void vqvdb_b200_destroy(vqvdb_b200_codec* codec) { delete codec; }

int vqvdb_b200_latent_shape(const vqvdb_b200_codec* codec, int64_t out_dhw[3]) {
	if (!codec || !out_dhw) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	out_dhw[0] = out_dhw[1] = out_dhw[2] = 4;
	return VQVDB_B200_OK;
}

int vqvdb_b200_in_channels(const vqvdb_b200_codec* codec) { return codec ? codec->channels : VQVDB_B200_ERR_INVALID_ARGUMENT; }
int vqvdb_b200_num_embeddings(const vqvdb_b200_codec* codec) { return codec ? codec->K : VQVDB_B200_ERR_INVALID_ARGUMENT; }

int vqvdb_b200_encode_device(vqvdb_b200_codec* c, const float* dev_leaves, int64_t n, uint8_t* dev_indices, void* stream) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || (n > 0 && (!dev_leaves || !dev_indices))) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "encode_device: bad arguments");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		launch_encode(*c, dev_leaves, n, dev_indices, (cudaStream_t)stream);
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_decode_device(vqvdb_b200_codec* c, const uint8_t* dev_indices, int64_t n, float* dev_voxels, void* stream) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || (n > 0 && (!dev_indices || !dev_voxels))) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "decode_device: bad arguments");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		launch_decode(*c, dev_indices, n, dev_voxels, (cudaStream_t)stream);
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_synchronize(vqvdb_b200_codec* c) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		CUDA_TRY(cudaStreamSynchronize(c->compute));
		if (c->staging_ready)
			for (auto& s : c->slots) CUDA_TRY(cudaStreamSynchronize(s.stream));
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

// Host-pointer encode: chunks of `chunk` leaves rotate through kSlots {stream, device buffers, pinned
// staging}; chunk i+1's staging copy and H2D overlap chunk i's kernel and chunk i-1's D2H.
int vqvdb_b200_encode(vqvdb_b200_codec* c, const float* host_leaves, int64_t n, uint8_t* host_indices) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || (n > 0 && (!host_leaves || !host_indices))) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "encode: bad arguments");
	if (n == 0) return VQVDB_B200_OK;
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		ensure_staging(*c);
		const size_t leaf_elems = (size_t)c->channels * 512;
		if (!c->generic && n <= std::min<int64_t>(kZeroCopyLeaves, c->chunk)) {  // small call: no copy engines (see kZeroCopyLeaves)
			Slot& s = c->slots[0];
			const float* src = device_alias(host_leaves, 16);  // null for pageable memory
			if (!src) {
				std::memcpy(s.h_vox, host_leaves, (size_t)n * leaf_elems * sizeof(float));
				src = device_alias(static_cast<const float*>(s.h_vox), 16);
			}
			uint8_t* dst = device_alias(host_indices, 1);
			const bool staged_out = dst == nullptr;
			if (staged_out) dst = device_alias(s.h_idx, 1);
			if (src && dst) {
				launch_encode(*c, src, n, dst, s.stream, 0);
				CUDA_TRY(cudaStreamSynchronize(s.stream));
				if (staged_out) std::memcpy(host_indices, s.h_idx, (size_t)n * 64);
				return VQVDB_B200_OK;
			}
		}
		const bool in_direct = is_pinned_host(host_leaves), out_direct = is_pinned_host(host_indices);
		int64_t done = 0;
		for (int i = 0; done < n; ++i) {
			Slot& s = c->slots[i % kSlots];
			retire<uint8_t>(*c->copier, s, host_indices, 64, s.h_idx, out_direct);
			const int64_t cnt = std::min<int64_t>(c->chunk, n - done);
			const float* src = host_leaves + (size_t)done * leaf_elems;
			if (!in_direct) {
				c->copier->copy(s.h_vox, src, (size_t)cnt * leaf_elems * sizeof(float));
				src = s.h_vox;
			}
			CUDA_TRY(cudaMemcpyAsync(s.d_vox, src, (size_t)cnt * leaf_elems * sizeof(float), cudaMemcpyHostToDevice, s.stream));
			launch_encode(*c, s.d_vox, cnt, s.d_idx, s.stream, i % kSlots);
			uint8_t* dst = out_direct ? host_indices + (size_t)done * 64 : s.h_idx;
			CUDA_TRY(cudaMemcpyAsync(dst, s.d_idx, (size_t)cnt * 64, cudaMemcpyDeviceToHost, s.stream));
			CUDA_TRY(cudaEventRecord(s.done, s.stream));
			s.pending_first = done;
			s.pending_count = cnt;
			done += cnt;
		}
		for (auto& s : c->slots) retire<uint8_t>(*c->copier, s, host_indices, 64, s.h_idx, out_direct);
	} catch (const std::exception& e) {
		for (auto& s : c->slots) s.pending_first = -1;
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_decode(vqvdb_b200_codec* c, const uint8_t* host_indices, int64_t n, float* host_voxels) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || (n > 0 && (!host_indices || !host_voxels))) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "decode: bad arguments");
	if (n == 0) return VQVDB_B200_OK;
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		ensure_staging(*c);
		const size_t leaf_elems = (size_t)c->channels * 512;
		if (!c->generic && n <= std::min<int64_t>(kZeroCopyLeaves, c->chunk)) {  // small call: no copy engines (see kZeroCopyLeaves)
			Slot& s = c->slots[0];
			const uint8_t* src = device_alias(host_indices, 4);  // null for pageable memory
			if (!src) {
				std::memcpy(s.h_idx, host_indices, (size_t)n * 64);
				src = device_alias(static_cast<const uint8_t*>(s.h_idx), 4);
			}
			float* dst = device_alias(host_voxels, 16);
			const bool staged_out = dst == nullptr;
			if (staged_out) dst = device_alias(s.h_vox, 16);
			if (src && dst) {
				launch_decode(*c, src, n, dst, s.stream, 0);
				CUDA_TRY(cudaStreamSynchronize(s.stream));
				if (staged_out) std::memcpy(host_voxels, s.h_vox, (size_t)n * leaf_elems * sizeof(float));
				return VQVDB_B200_OK;
			}
		}
		const bool in_direct = is_pinned_host(host_indices), out_direct = is_pinned_host(host_voxels);
		int64_t done = 0;
		for (int i = 0; done < n; ++i) {
			Slot& s = c->slots[i % kSlots];
			retire<float>(*c->copier, s, host_voxels, leaf_elems, s.h_vox, out_direct);
			const int64_t cnt = std::min<int64_t>(c->chunk, n - done);
			const uint8_t* src = host_indices + (size_t)done * 64;
			if (!in_direct) {
				c->copier->copy(s.h_idx, src, (size_t)cnt * 64);
				src = s.h_idx;
			}
			CUDA_TRY(cudaMemcpyAsync(s.d_idx, src, (size_t)cnt * 64, cudaMemcpyHostToDevice, s.stream));
			launch_decode(*c, s.d_idx, cnt, s.d_vox, s.stream, i % kSlots);
			float* dst = out_direct ? host_voxels + (size_t)done * leaf_elems : s.h_vox;
			CUDA_TRY(cudaMemcpyAsync(dst, s.d_vox, (size_t)cnt * leaf_elems * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
			CUDA_TRY(cudaEventRecord(s.done, s.stream));
			s.pending_first = done;
			s.pending_count = cnt;
			done += cnt;
		}
		for (auto& s : c->slots) retire<float>(*c->copier, s, host_voxels, leaf_elems, s.h_vox, out_direct);
	} catch (const std::exception& e) {
		for (auto& s : c->slots) s.pending_first = -1;
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_debug_decode_tap(vqvdb_b200_codec* c, const uint8_t* dev_indices, int64_t n, int stage, float* dev_tap,
                                float* dev_voxels, void* stream) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || stage < 0 || (stage > 2 && stage != 100) || (n > 0 && (!dev_indices || !dev_tap || !dev_voxels)))
		return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "debug_decode_tap: bad arguments");
	if (c->generic && !c->dec128) return fail(c, VQVDB_B200_ERR_UNSUPPORTED, "debug_decode_tap: no tensor-core decoder for this model");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		if (c->generic) CUDA_TRY(vqvdb::launch_decode_tc128(c->dec128_w, dev_indices, n, dev_voxels, c->num_sms, (cudaStream_t)stream, stage, dev_tap));
		else CUDA_TRY(vqvdb::launch_decode_tc(c->dec_mma, dev_indices, n, dev_voxels, c->num_sms, (cudaStream_t)stream, stage, dev_tap));
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_debug_encode_tap(vqvdb_b200_codec* c, const float* dev_leaves, int64_t n, int stage, float* dev_tap,
                                uint8_t* dev_indices, void* stream) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (n < 0 || stage < 0 || (stage > 7 && stage != 100) || (n > 0 && (!dev_leaves || !dev_tap || !dev_indices)))
		return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "debug_encode_tap: bad arguments");
	if (c->generic && !c->enc128) return fail(c, VQVDB_B200_ERR_UNSUPPORTED, "debug_encode_tap: no tensor-core encoder for this model");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		if (c->generic) {  // vec3: stages 0..3 = res_stack.0, res_stack.1, attention, proj, each [leaf][128][64]
			// ... 4, 5 = pre (GroupNorm + ReLU), the 8^3 residual block, each [leaf][64][512]; 6 = down1 [leaf][128][64]
			if ((stage > 6 && stage != 100) || n > c->enc128_batch) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "debug_encode_tap: bad stage or too many leaves");
			float* scratch = c->enc128_scratch + (size_t)kSlots * vqvdb::encode_tc128_front_scratch_floats(c->num_sms);
			float* y = c->enc128_y + (size_t)kSlots * c->enc128_batch * 8192;
			const cudaStream_t st = (cudaStream_t)stream;
			if (kEnc128GenericFront) {
				float* gscratch = c->gen_scratch + (size_t)kSlots * vqvdb::generic_scratch_floats(c->gen_grid);
				CUDA_TRY(vqvdb::launch_encode_generic_front(c->gen, dev_leaves, n, y, gscratch, c->gen_grid, st));
			} else CUDA_TRY(vqvdb::launch_encode_tc128_front(c->enc128_front, c->enc128_pre, dev_leaves, n, y, scratch, c->num_sms, st, stage == 100 ? 100 : stage >= 4 ? stage - 4 : -1, dev_tap));
			if (stage == 6) CUDA_TRY(cudaMemcpyAsync(dev_tap, y, (size_t)n * 8192 * sizeof(float), cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(vqvdb::launch_encode_tc128_back(c->enc128_back, y, n, dev_indices, c->num_sms, st, stage <= 3 || stage == 100 ? stage : -1, dev_tap));
			return VQVDB_B200_OK;
		}
		CUDA_TRY(vqvdb::launch_encode_tc(c->enc, c->enc_tc, dev_leaves, n, dev_indices, c->num_sms, (cudaStream_t)stream, stage, dev_tap));
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_peer_buffer_create(vqvdb_b200_codec* c, uint64_t bytes, void** dev_ptr_out, unsigned char handle_out[64]) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (!bytes || !dev_ptr_out || !handle_out) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "peer_buffer_create: bad arguments");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		void* p = nullptr;
		CUDA_TRY(cudaMalloc(&p, bytes));
		cudaIpcMemHandle_t h;
		cudaError_t e = cudaIpcGetMemHandle(&h, p);
		if (e != cudaSuccess) {
			cudaFree(p);
			throw CudaError(e, "cudaIpcGetMemHandle");
		}
		std::memcpy(handle_out, &h, 64);
		*dev_ptr_out = p;
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_peer_buffer_open(vqvdb_b200_codec* c, const unsigned char handle[64], void** dev_ptr_out) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (!handle || !dev_ptr_out) return fail(c, VQVDB_B200_ERR_INVALID_ARGUMENT, "peer_buffer_open: bad arguments");
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		cudaIpcMemHandle_t h;
		std::memcpy(&h, handle, 64);
		void* p = nullptr;
		CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		*dev_ptr_out = p;
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_peer_buffer_close(vqvdb_b200_codec* c, void* dev_ptr, int opened) {
	if (!c) return VQVDB_B200_ERR_INVALID_ARGUMENT;
	if (!dev_ptr) return VQVDB_B200_OK;
	try {
		CUDA_TRY(cudaSetDevice(c->device));
		if (opened) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
		else CUDA_TRY(cudaFree(dev_ptr));
	} catch (const std::exception& e) {
		return translate(c, e);
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_convert_onnx(const char* enc_path, const char* dec_path, const char* out_path) {
	if (!enc_path || !dec_path || !out_path) return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "convert_onnx: bad arguments");
	try {
		const auto blob = vqvdb::onnx_to_pack(vqvdb::read_file_bytes(enc_path), vqvdb::read_file_bytes(dec_path));
		WeightPack check;
		check.parse(blob.data(), blob.size());
		std::ofstream f(out_path, std::ios::binary);
		if (!f || !f.write(reinterpret_cast<const char*>(blob.data()), (std::streamsize)blob.size())) throw std::runtime_error(std::string("cannot write ") + out_path);
	} catch (const std::exception& e) {
		return fail(nullptr, VQVDB_B200_ERR_BAD_WEIGHTS, e.what());
	}
	return VQVDB_B200_OK;
}

uint64_t vqvdb_b200_kernel_launches(const vqvdb_b200_codec* c) { return c ? c->launches.load() : 0; }
int vqvdb_b200_debug_fold_decoder_tail(const char* weights_path, float* weights_out, float* bias_out) {
	if (!weights_out || !bias_out) return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "debug_fold_decoder_tail: null output");
	try {
		WeightPack pack;
		if (weights_path && weights_path[0]) pack.load_file(weights_path);
		else pack.parse(vqvdb::vqvdb_b200_embedded_pack, (size_t)(vqvdb::vqvdb_b200_embedded_pack_end - vqvdb::vqvdb_b200_embedded_pack));
		std::vector<float> wg, bg;
		vqvdb::build_decoder_fold(pack, wg, bg);
		std::memcpy(weights_out, wg.data(), wg.size() * sizeof(float));
		std::memcpy(bias_out, bg.data(), bg.size() * sizeof(float));
	} catch (const std::exception& e) {
		return fail(nullptr, VQVDB_B200_ERR_BAD_WEIGHTS, e.what());
	}
	return VQVDB_B200_OK;
}

int vqvdb_b200_debug_fold_encoder_vq(const char* weights_path, float* m_out, float* esq_out, float* norm_out) {
	if (!m_out || !esq_out || !norm_out) return fail(nullptr, VQVDB_B200_ERR_INVALID_ARGUMENT, "debug_fold_encoder_vq: null output");
	try {
		WeightPack pack;
		if (weights_path && weights_path[0]) pack.load_file(weights_path);
		else pack.parse(vqvdb::vqvdb_b200_embedded_pack, (size_t)(vqvdb::vqvdb_b200_embedded_pack_end - vqvdb::vqvdb_b200_embedded_pack));
		std::vector<float> m, esq, mno;
		if (vqvdb::encoder128_supports(pack)) {  // vec3 model: M [256][128]; esq and the bound's three constants from the parameter block
			m = vqvdb::build_encoder128_vq_fold(pack);
			const std::vector<float> par = vqvdb::build_encoder128_back_params(pack);
			esq.assign(par.begin() + vqvdb::par128e::vq_esq2, par.begin() + vqvdb::par128e::vq_esq2 + 256);
			mno.assign(257, 0.f);
			for (int i = 0; i < 3; ++i) mno[i] = par[vqvdb::par128e::vq_const + i];
		} else
			vqvdb::build_encoder_vq_fold(pack, m, esq, mno);
		std::memcpy(m_out, m.data(), m.size() * sizeof(float));
		std::memcpy(esq_out, esq.data(), esq.size() * sizeof(float));
		std::memcpy(norm_out, mno.data(), mno.size() * sizeof(float));
	} catch (const std::exception& e) {
		return fail(nullptr, VQVDB_B200_ERR_BAD_WEIGHTS, e.what());
	}
	return VQVDB_B200_OK;
}

const char* vqvdb_b200_decode_path(const vqvdb_b200_codec* c) { return c ? c->decode_path.c_str() : ""; }
const char* vqvdb_b200_encode_path(const vqvdb_b200_codec* c) { return c ? c->encode_path.c_str() : ""; }

const char* vqvdb_b200_last_error(const vqvdb_b200_codec* c) { return c ? c->last_error.c_str() : g_create_error.c_str(); }

}  // extern "C"
