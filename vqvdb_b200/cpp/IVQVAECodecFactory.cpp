// IVQVAECodec::create for this repository's build.  Same contract as the reference's factory
// (src/core/IVQVAECodec.cpp:76-110): swallow, log to stderr, return null.  In the reference tree this file is
// NOT added — the reference keeps its own factory and gains one `case BackendType::B200`.
#include <iostream>
#include <stdexcept>

#include "B200Backend.hpp"

std::unique_ptr<IVQVAECodec> IVQVAECodec::create(const CodecConfig& config, BackendType type) {
	try {
		if (type == BackendType::B200) return std::make_unique<B200Backend>(config);
		throw std::runtime_error("Requested backend type is not available or disabled in the build configuration.");
	} catch (const std::exception& e) {
		std::cerr << "Failed to create VQ-VAE backend: " << e.what() << std::endl;
		return nullptr;
	}
}
