// LayerRegistry and the DEFINE_LAYER_CREATOR / REGISTER_LAYER_CREATOR plugin macros — same API as the
// reference (/root/reference/src/layer_factory.h:31-90).
#pragma once

#include <map>
#include <string>
#include <vector>

#include "layer.h"

namespace feather {
inline namespace b200 {  // ABI tag: keeps these symbols apart from the reference build when both are loaded

class LayerRegistry {
public:
    typedef Layer* (*Creator)(RuntimeParameter<float>*);
    typedef std::map<std::string, Creator> CreatorRegistry;

    static CreatorRegistry& Registry() {
        static CreatorRegistry* g_registry_ = new CreatorRegistry();
        return *g_registry_;
    }
    static void AddCreator(const std::string& type, Creator creator) { Registry()[type] = creator; }
    static Layer* CreateLayer(std::string type, RuntimeParameter<float>* rt_param) {
        CreatorRegistry& registry = Registry();
        CreatorRegistry::iterator it = registry.find(type);
        if (it != registry.end()) return it->second(rt_param);
        fprintf(stderr, "Layer type %s is not supported in FeatherCNN...Aborting\n", type.c_str());
        return NULL;
    }

private:
    LayerRegistry() {}
};

class LayerRegisterer {
public:
    LayerRegisterer(const std::string& type, Layer* (*creator)(RuntimeParameter<float>*)) {
        LayerRegistry::AddCreator(type, creator);
    }
};

void register_layer_creators();

#define DEFINE_LAYER_CREATOR(feather_layer_name)                                       \
    static Layer* GetLayer##feather_layer_name(RuntimeParameter<float>* rt_param) {    \
        return (Layer*)new feather_layer_name##Layer(rt_param);                        \
    }

#define REGISTER_LAYER_CREATOR(ncnn_type_name, feather_layer_name) \
    static LayerRegisterer g_creator_f_##ncnn_type_name(#ncnn_type_name, GetLayer##feather_layer_name);

}  // inline namespace b200
}  // namespace feather
