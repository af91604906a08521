// Reads a reference TorchScript model (.pt) with libtorch and hands it to the engine through the C ABI.
// libtorch is used for nothing else: hyper-parameters come from the module's exported getters exactly as
// network/network.cpp:21-41 reads them, tensors from named_parameters() + named_buffers().
#pragma once
#include "../../include/mz_b200.h"
#include <string>

namespace mzhost {

struct NetInfo {
    std::string game_name, type_name;
    mz_net_dims dims{};
};

// reads only the hyper-parameters (no engine needed): used to create the engines for the right game
bool readNetInfo(const std::string& path, NetInfo& info, std::string& error);
// configure + set tensors + finalize on `engine`
bool loadNetwork(const std::string& path, mz_engine* engine, std::string& error);
// ranks that receive the packed blob instead of reading the file
bool configureEmpty(const NetInfo& info, mz_engine* engine, std::string& error);

} // namespace mzhost
