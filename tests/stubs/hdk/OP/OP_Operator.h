#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <PRM/PRM_Include.h>
class OP_Network;
class OP_Operator;
class CH_LocalVariable;
class OP_Context {
   public:
	fpreal getTime() const;
};
class OP_Node {
   public:
	virtual ~OP_Node();
};
using OP_Constructor = OP_Node* (*)(OP_Network*, const char*, OP_Operator*);
enum { OP_FLAG_GENERATOR = 0x04 };
class OP_Operator {
   public:
	OP_Operator(const char* name, const char* english, OP_Constructor construct, PRM_Template* templates, unsigned min_sources,
	            unsigned max_sources = 9999, CH_LocalVariable* variables = nullptr, unsigned flags = 0);
};
