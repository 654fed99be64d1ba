#pragma once  // TEST INFRASTRUCTURE (see SYS/SYS_Types.h)
#include <GA/GA_Iterator.h>
#include <openvdb/openvdb.h>
extern const GA_PrimitiveTypeId GEO_PRIMVDB;
class GEO_Primitive {
   public:
	virtual ~GEO_Primitive();
	const GA_PrimitiveTypeId& getTypeId() const;
};
class GEO_PrimVDB : public GEO_Primitive {
   public:
	openvdb::GridBase::ConstPtr getConstGridPtr() const;
	openvdb::GridBase::Ptr getGridPtr();
	const char* getGridName() const;
};
