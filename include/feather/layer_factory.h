// Layer registry and plugin macros.
//
// Public surface kept from the reference (/root/reference/src/layer_factory.h:31-90) because user code is written
// against it:  LayerRegistry::CreateLayer / AddCreator / Registry,  LayerRegisterer,  register_layer_creators(),
// DEFINE_LAYER_CREATOR(X)  and  REGISTER_LAYER_CREATOR(ncnnType, X).  The registry itself lives in
// libfeather_b200.so (layer_factory.cpp), not in an inline function, so there is exactly one table per process even
// when several shared objects include this header.
#pragma once

#include <map>
#include <string>

#include "layer.h"

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class LayerRegistry {
public:
    using Creator = Layer* (*)(RuntimeParameter<float>*);
    using CreatorRegistry = std::map<std::string, Creator>;

    // type name -> factory; filled by register_layer_creators() and by REGISTER_LAYER_CREATOR in user code
    static CreatorRegistry& Registry();
    static void AddCreator(const std::string& type, Creator creator);
    // NULL (and a message on stderr) for an unknown type, like the reference
    static Layer* CreateLayer(std::string type, RuntimeParameter<float>* rt_param);

private:
    LayerRegistry();  // static-only
};

// Registers at static-initialisation time; what REGISTER_LAYER_CREATOR instantiates.
class LayerRegisterer {
public:
    LayerRegisterer(const std::string& type, LayerRegistry::Creator creator) { LayerRegistry::AddCreator(type, creator); }
};

// The 13 built-in ncnn layer types.
void register_layer_creators();

}  // inline namespace b200
}  // namespace feather

// Plugin macros for a class named <X>Layer with a constructor taking RuntimeParameter<float>*.
#define DEFINE_LAYER_CREATOR(X) \
    static ::feather::Layer* GetLayer##X(::RuntimeParameter<float>* rt_param) { return new X##Layer(rt_param); }

#define REGISTER_LAYER_CREATOR(ncnn_type_name, X) \
    static ::feather::LayerRegisterer g_creator_f_##ncnn_type_name(#ncnn_type_name, GetLayer##X);
