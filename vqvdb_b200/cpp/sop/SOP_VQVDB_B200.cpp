// Houdini SOP shim for the B200 backend: the two nodes of the reference's plugin —
// `vqvdb_encoder` (src/SOP/SOP_VQVDB_Encoder.{hpp,cpp}) and `vqvdb_decoder` (src/SOP/SOP_VQVDB_Decoder.{hpp,cpp}),
// registered by newSopOperator (src/SOP/SOP_Register.cpp:15-18) — with the same operator names, parameter names and
// cook behaviour, created on BackendType::B200 instead of BackendType::ONNX (SOP_VQVDB_Encoder.cpp:67,
// SOP_VQVDB_Decoder.cpp:62).
//
// Compile-guarded: the Houdini Development Kit is not present in the image this repository is built in, so `make`
// never compiles this file.  tests/test_sop_shim.py syntax-checks it against declaration-only stand-ins for the few
// HDK classes it touches (tests/stubs/hdk/) and the OpenVDB stand-in of tests/stubs/openvdb/; it has NOT been
// compiled against a real HDK or loaded into Houdini.  It is written with the classic SOP_Node::cookMySop entry
// point and a PRM_Template table, which need no generated *.proto.h (the reference's verbs do).
//
// What differs from the reference's nodes, on purpose:
//   * `vdbname` is honoured.  The reference declares the parm (SOP_VQVDB_Encoder.cpp:20-25, SOP_VQVDB_Decoder.cpp:21-26)
//     and never reads it; here it filters the VDB primitives by grid name (empty or "*" = every VDB, the
//     reference's effective behaviour).
//   * interrupts are polled: VQVAECodec.hpp:38,47 of the reference document a `boss` argument "for progress and
//     cancellation" that does not exist; here UT_AutoInterrupt is checked between backend calls.
//   * `batchsize` 0 (the default here) hands a whole grid to the backend per call — it pipelines internally in 16 K-leaf
//     chunks — and any other value is used as given, like the reference's (default 64, SOP_VQVDB_Encoder.cpp:36).
#if defined(VQVDB_B200_WITH_HDK) || (defined(__has_include) && __has_include(<SOP/SOP_Node.h>))

#include <GA/GA_Iterator.h>
#include <GEO/GEO_PrimVDB.h>
#include <GU/GU_Detail.h>
#include <GU/GU_PrimVDB.h>
#include <OP/OP_AutoLockInputs.h>
#include <OP/OP_Operator.h>
#include <OP/OP_OperatorTable.h>
#include <PRM/PRM_Include.h>
#include <SOP/SOP_Node.h>
#include <UT/UT_DSOVersion.h>
#include <UT/UT_Interrupt.h>
#include <UT/UT_String.h>

#include <filesystem>
#include <memory>
#include <string>
#include <vector>

#define VQVDB_B200_WITH_OPENVDB 1
#include "../B200Backend.hpp"
#include "../openvdb_adapter.hpp"

namespace {

// One codec per node, created on first cook and kept for the node's life (the reference's SOP_NodeCache does the same:
// SOP_VQVDB_Encoder.cpp:57-81).  create() never throws; a null backend is reported as a cook error.
std::unique_ptr<VQVAECodec> makeCodec(std::string& why) {
	CodecConfig config;
	config.device = CodecConfig::Device::CUDA;
	config.source = EmbeddedModel{};
	try {
		std::unique_ptr<IVQVAECodec> backend = IVQVAECodec::create(config, BackendType::B200);
		if (!backend) {
			why = "Failed to initialize VQ-VDB codec backend (no sm_100 device, or libvqvdb_b200.so missing; there is no CPU fallback).";
			return nullptr;
		}
		return std::make_unique<VQVAECodec>(std::move(backend));
	} catch (const std::exception& e) {
		why = e.what();
		return nullptr;
	}
}

bool nameSelected(const UT_String& pattern, const std::string& gridName) {
	if (!pattern.isstring() || pattern == "*") return true;
	return UT_String(gridName.c_str()).multiMatch(pattern);
}

PRM_Name kVdbName("vdbname", "VDB Grid Name");
PRM_Default kVdbNameDefault(0, "density");
PRM_Name kOutputPath("outputpath", "Output File (.vqvdb)");
PRM_Name kInputFile("inputfile", "Input File (.vqvdb)");
PRM_Name kBatchSize("batchsize", "GPU Batch Size (0 = whole grid)");
PRM_Default kBatchDefault(0);
PRM_Range kEncBatchRange(PRM_RANGE_RESTRICTED, 0, PRM_RANGE_UI, 1024);
PRM_Range kDecBatchRange(PRM_RANGE_RESTRICTED, 0, PRM_RANGE_UI, 8192);
PRM_Name kExecute("execute", "Encode and Save to Disk");

}  // namespace

// ------------------------------------------------------------------------------------------------ encoder
class SOP_VQVDB_B200_Encoder final : public SOP_Node {
   public:
	static OP_Node* myConstructor(OP_Network* net, const char* name, OP_Operator* op) { return new SOP_VQVDB_B200_Encoder(net, name, op); }
	static PRM_Template myTemplateList[];
	const char* inputLabel(unsigned idx) const override { return idx == 0 ? "Input Grids" : "Sourcing Grids"; }

   protected:
	SOP_VQVDB_B200_Encoder(OP_Network* net, const char* name, OP_Operator* op) : SOP_Node(net, name, op) {}
	OP_ERROR cookMySop(OP_Context& context) override;

   private:
	std::unique_ptr<VQVAECodec> codec_;
};

PRM_Template SOP_VQVDB_B200_Encoder::myTemplateList[] = {
    PRM_Template(PRM_STRING, 1, &kVdbName, &kVdbNameDefault),
    PRM_Template(PRM_FILE, 1, &kOutputPath),
    PRM_Template(PRM_INT, 1, &kBatchSize, &kBatchDefault, nullptr, &kEncBatchRange),
    PRM_Template(PRM_TOGGLE, 1, &kExecute, PRMzeroDefaults),
    PRM_Template(),
};

OP_ERROR SOP_VQVDB_B200_Encoder::cookMySop(OP_Context& context) {
	const fpreal now = context.getTime();
	OP_AutoLockInputs inputs(this);
	if (inputs.lock(context) >= UT_ERROR_ABORT) return error();
	duplicateSource(0, context);  // the node passes its input through (the reference cooks with COOK_DUPLICATE)
	if (evalInt("execute", 0, now) == 0) return error();
	setInt("execute", 0, now, 0);  // always reset the toggle, even if encoding fails below (SOP_VQVDB_Encoder.cpp:90-91)

	std::string why;
	if (!codec_ && !(codec_ = makeCodec(why))) {
		addError(SOP_MESSAGE, why.c_str());
		return error();
	}
	UT_String pattern, outPath;
	evalString(pattern, "vdbname", 0, now);
	evalString(outPath, "outputpath", 0, now);
	if (!outPath.isstring()) {
		addError(SOP_MESSAGE, "Output path must be specified.");
		return error();
	}

	std::vector<openvdb::FloatGrid::ConstPtr> grids;
	for (GA_Iterator it(gdp->getPrimitiveRange()); !it.atEnd(); ++it) {
		const GEO_Primitive* prim = gdp->getGEOPrimitive(*it);
		if (!prim || prim->getTypeId() != GEO_PRIMVDB) continue;
		const auto* vdb = static_cast<const GEO_PrimVDB*>(prim);
		openvdb::GridBase::ConstPtr grid = vdb->getConstGridPtr();
		if (!grid || !nameSelected(pattern, grid->getName())) continue;
		if (auto fg = openvdb::gridConstPtrCast<openvdb::FloatGrid>(grid)) grids.push_back(fg);
		else {  // the model is trained on scalar float leaves; the reference rejects other grids the same way (:123-128)
			addError(SOP_MESSAGE, ("Skipped non-float grid: " + grid->getName()).c_str());
			return error();
		}
	}
	if (grids.empty()) {
		addError(SOP_MESSAGE, "Failed to load VDB grid from input.");
		return error();
	}
	try {
		UT_AutoInterrupt boss("Encoding VDB leaves (VQ-VAE, B200)");
		size_t batch = (size_t)evalInt("batchsize", 0, now);
		if (batch == 0) batch = VQVAECodec::kInterruptibleBatch;
		vqvdb_openvdb::compress(*codec_, grids, std::filesystem::path(outPath.toStdString()), batch, [&boss] { return boss.wasInterrupted(); });
		addMessage(SOP_MESSAGE, ("Successfully saved to " + outPath.toStdString()).c_str());
	} catch (const std::exception& e) {
		addError(SOP_MESSAGE, e.what());
	}
	return error();
}

// ------------------------------------------------------------------------------------------------ decoder
class SOP_VQVDB_B200_Decoder final : public SOP_Node {
   public:
	static OP_Node* myConstructor(OP_Network* net, const char* name, OP_Operator* op) { return new SOP_VQVDB_B200_Decoder(net, name, op); }
	static PRM_Template myTemplateList[];

   protected:
	SOP_VQVDB_B200_Decoder(OP_Network* net, const char* name, OP_Operator* op) : SOP_Node(net, name, op) {}
	OP_ERROR cookMySop(OP_Context& context) override;

   private:
	std::unique_ptr<VQVAECodec> codec_;
};

PRM_Template SOP_VQVDB_B200_Decoder::myTemplateList[] = {
    PRM_Template(PRM_STRING, 1, &kVdbName, &kVdbNameDefault),
    PRM_Template(PRM_FILE, 1, &kInputFile),
    PRM_Template(PRM_INT, 1, &kBatchSize, &kBatchDefault, nullptr, &kDecBatchRange),
    PRM_Template(),
};

OP_ERROR SOP_VQVDB_B200_Decoder::cookMySop(OP_Context& context) {
	const fpreal now = context.getTime();
	std::string why;
	if (!codec_ && !(codec_ = makeCodec(why))) {
		addError(SOP_MESSAGE, why.c_str());
		return error();
	}
	UT_String pattern, inFile;
	evalString(pattern, "vdbname", 0, now);
	evalString(inFile, "inputfile", 0, now);
	if (!inFile.isstring()) return error();  // no file specified: nothing to do (SOP_VQVDB_Decoder.cpp:87-90)
	const std::filesystem::path inPath(inFile.toStdString());
	if (!std::filesystem::exists(inPath)) {
		addError(SOP_MESSAGE, "Input file does not exist.");
		return error();
	}
	std::vector<openvdb::FloatGrid::Ptr> grids;
	try {
		UT_AutoInterrupt boss("Decoding VDB leaves (VQ-VAE, B200)");
		size_t batch = (size_t)evalInt("batchsize", 0, now);
		if (batch == 0) batch = VQVAECodec::kInterruptibleBatch;
		grids = vqvdb_openvdb::decompress(*codec_, inPath, 0.0f, batch, [&boss] { return boss.wasInterrupted(); });
	} catch (const std::exception& e) {
		addError(SOP_MESSAGE, e.what());
		return error();
	}
	gdp->clearAndDestroy();  // a generator: its geometry is exactly the decoded grids (SOP_VQVDB_Decoder.cpp:109-114)
	for (const auto& grid : grids)
		if (nameSelected(pattern, grid->getName())) GU_PrimVDB::buildFromGrid(*gdp, grid, nullptr, grid->getName().c_str());
	return error();
}

// ------------------------------------------------------------------------------------------------ registration
// The single entry point Houdini calls for this DSO; same operator names as the reference's plugin, so scenes that
// reference `vqvdb_encoder` / `vqvdb_decoder` keep working (install one DSO or the other, not both).
void newSopOperator(OP_OperatorTable* table) {
	table->addOperator(new OP_Operator("vqvdb_encoder", "VQVDB Encoder", SOP_VQVDB_B200_Encoder::myConstructor,
	                                   SOP_VQVDB_B200_Encoder::myTemplateList, 1, 1, nullptr));
	table->addOperator(new OP_Operator("vqvdb_decoder", "VQVDB Decoder", SOP_VQVDB_B200_Decoder::myConstructor,
	                                   SOP_VQVDB_B200_Decoder::myTemplateList, 0, 0, nullptr, OP_FLAG_GENERATOR));
}

#endif  // HDK available
