#include "net_loader.h"
#include <torch/script.h>
#include <vector>

namespace mzhost {

namespace {
bool fill(torch::jit::script::Module& m, NetInfo& info)
{
    auto geti = [&](const char* name) { return static_cast<int32_t>(m.get_method(name)({}).toInt()); };
    info.game_name = m.get_method("get_game_name")({}).toString()->string();
    info.type_name = m.get_method("get_type_name")({}).toString()->string();
    info.dims.num_input_channels = geti("get_num_input_channels");
    info.dims.input_height = geti("get_input_channel_height");
    info.dims.input_width = geti("get_input_channel_width");
    info.dims.num_hidden_channels = geti("get_num_hidden_channels");
    info.dims.num_blocks = geti("get_num_blocks");
    info.dims.action_size = geti("get_action_size");
    info.dims.num_value_hidden_channels = geti("get_num_value_hidden_channels");
    info.dims.discrete_value_size = geti("get_discrete_value_size");
    info.dims.is_muzero = (info.type_name == "muzero" || info.type_name == "muzero_atari");
    info.dims.num_action_feature_channels = (info.dims.is_muzero ? geti("get_num_action_feature_channels") : 0); // network/muzero_network.h:51
    return true;
}
} // namespace

bool readNetInfo(const std::string& path, NetInfo& info, std::string& error)
{
    try {
        torch::jit::script::Module m = torch::jit::load(path, torch::kCPU);
        return fill(m, info);
    } catch (const std::exception& e) {
        error = e.what();
        return false;
    }
}

bool loadNetwork(const std::string& path, mz_engine* engine, std::string& error)
{
    try {
        torch::jit::script::Module m = torch::jit::load(path, torch::kCPU);
        m.eval();
        NetInfo info;
        fill(m, info);
        if (info.type_name != "alphazero" && info.type_name != "muzero" && info.type_name != "muzero_atari") {
            error = "network type '" + info.type_name + "' is not supported by this engine";
            return false;
        }
        if (mz_net_configure(engine, &info.dims) != MZ_OK) {
            error = mz_last_error();
            return false;
        }
        auto push = [&](const std::string& name, const at::Tensor& t) {
            if (name.size() > 19 && name.compare(name.size() - 19, 19, "num_batches_tracked") == 0) { return true; }
            at::Tensor f = t.detach().to(torch::kCPU, torch::kFloat32).contiguous();
            return mz_net_set_tensor(engine, name.c_str(), f.data_ptr<float>(), f.numel()) == MZ_OK;
        };
        for (const auto& p : m.named_parameters()) {
            if (!push(p.name, p.value)) {
                error = mz_last_error();
                return false;
            }
        }
        for (const auto& b : m.named_buffers()) {
            if (!push(b.name, b.value)) {
                error = mz_last_error();
                return false;
            }
        }
        if (mz_net_finalize(engine) != MZ_OK) {
            error = mz_last_error();
            return false;
        }
        return true;
    } catch (const std::exception& e) {
        error = e.what();
        return false;
    }
}

bool configureEmpty(const NetInfo& info, mz_engine* engine, std::string& error)
{
    if (mz_net_configure(engine, &info.dims) != MZ_OK || mz_net_finalize_empty(engine) != MZ_OK) {
        error = mz_last_error();
        return false;
    }
    return true;
}

} // namespace mzhost
