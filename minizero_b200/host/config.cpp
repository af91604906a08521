#include "config.h"
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace mzhost {

namespace {
struct Def {
    const char* key;
    char type;
    const char* value;
};
// keys and defaults of config/configuration.cpp:6-90 (the reference refuses unknown keys, so all are registered)
const Def kDefs[] = {
    {"program_seed", 'i', "0"}, {"program_auto_seed", 'b', "false"}, {"program_quiet", 'b', "false"}, {"program_use_color_message", 'b', "true"},
    {"actor_num_simulation", 'i', "50"}, {"actor_mcts_puct_base", 'f', "19652"}, {"actor_mcts_puct_init", 'f', "1.25"},
    {"actor_mcts_reward_discount", 'f', "1"}, {"actor_mcts_value_rescale", 'b', "false"}, {"actor_mcts_think_batch_size", 'i', "1"},
    {"actor_mcts_think_time_limit", 'f', "0"}, {"actor_select_action_by_count", 'b', "false"}, {"actor_select_action_by_softmax_count", 'b', "true"},
    {"actor_select_action_softmax_temperature", 'f', "1"}, {"actor_select_action_softmax_temperature_decay", 'b', "false"},
    {"actor_use_random_rotation_features", 'b', "true"}, {"actor_use_dirichlet_noise", 'b', "true"}, {"actor_dirichlet_noise_alpha", 'f', "0.03"},
    {"actor_dirichlet_noise_epsilon", 'f', "0.25"}, {"actor_use_gumbel", 'b', "false"}, {"actor_use_gumbel_noise", 'b', "false"},
    {"actor_gumbel_sample_size", 'i', "16"}, {"actor_gumbel_sigma_visit_c", 'f', "50"}, {"actor_gumbel_sigma_scale_c", 'f', "1"},
    {"actor_resign_threshold", 'f', "-0.9"},
    {"zero_num_threads", 'i', "4"}, {"zero_num_parallel_games", 'i', "32"}, {"zero_server_port", 'i', "9999"}, {"zero_training_directory", 's', ""},
    {"zero_num_games_per_iteration", 'i', "2000"}, {"zero_start_iteration", 'i', "0"}, {"zero_end_iteration", 'i', "100"}, {"zero_replay_buffer", 'i', "20"},
    {"zero_disable_resign_ratio", 'f', "0.1"}, {"zero_actor_intermediate_sequence_length", 'i', "0"}, {"zero_actor_ignored_command", 's', "reset_actors"},
    {"zero_server_accept_different_model_games", 'b', "true"}, {"zero_display_latest_games", 'i', "0"},
    {"learner_use_per", 'b', "false"}, {"learner_per_alpha", 'f', "1"}, {"learner_per_init_beta", 'f', "1"}, {"learner_per_beta_anneal", 'b', "true"},
    {"learner_training_step", 'i', "500"}, {"learner_training_display_step", 'i', "100"}, {"learner_batch_size", 'i', "1024"},
    {"learner_muzero_unrolling_step", 'i', "5"}, {"learner_n_step_return", 'i', "0"}, {"learner_optimizer", 's', "SGD"}, {"learner_learning_rate", 'f', "0.02"},
    {"learner_momentum", 'f', "0.9"}, {"learner_weight_decay", 'f', "0.0001"}, {"learner_value_loss_scale", 'f', "1"}, {"learner_num_thread", 'i', "8"},
    {"nn_file_name", 's', ""}, {"nn_num_blocks", 'i', "1"}, {"nn_num_hidden_channels", 'i', "256"}, {"nn_num_value_hidden_channels", 'i', "256"},
    {"nn_type_name", 's', "alphazero"},
    {"env_board_size", 'i', "0"}, {"env_atari_rom_dir", 's', "/opt/atari57/"}, {"env_atari_name", 's', "ms_pacman"}, {"env_conhex_use_swap_rule", 'b', "true"},
    {"env_go_komi", 'f', "7.5"}, {"env_go_ko_rule", 's', "positional"}, {"env_gomoku_rule", 's', "standard"}, {"env_gomoku_exactly_five_stones", 'b', "true"},
    {"env_havannah_use_swap_rule", 'b', "true"}, {"env_hex_use_swap_rule", 'b', "true"}, {"env_killallgo_ko_rule", 's', "positional"},
    {"env_killallgo_use_seki", 'b', "false"}, {"env_rubiks_scramble_rotate", 'i', "5"}, {"env_surakarta_no_capture_plies", 'i', "50"},
    {"env_tetris_block_puzzle_num_holding_block", 'i', "3"}, {"env_tetris_block_puzzle_num_preview_holding_block", 'i', "0"},
};

void trim(std::string& s)
{
    if (s.empty()) { return; }
    const size_t b = s.find_first_not_of(" \t");
    if (b == std::string::npos) {
        s.clear();
        return;
    }
    s.erase(0, b);
    s.erase(s.find_last_not_of(" \t\r") + 1);
}
} // namespace

Config::Config()
{
    for (const Def& d : kDefs) {
        values_[d.key] = d.value;
        types_[d.key] = d.type;
    }
}

bool Config::loadFromFile(const std::string& path)
{
    if (path.empty()) { return false; }
    std::ifstream file(path);
    if (file.fail()) { return false; }
    std::string line;
    while (std::getline(file, line)) {
        if (!setValue(line)) { return false; }
    }
    return true;
}

bool Config::loadFromString(const std::string& s)
{
    if (s.empty()) { return false; }
    std::istringstream iss(s);
    std::string line;
    while (std::getline(iss, line, ':')) {
        if (!setValue(line)) { return false; }
    }
    return true;
}

bool Config::setValue(std::string line)
{
    if (line.empty() || line[0] == '#') { return true; }
    std::string key = line.substr(0, line.find("="));
    std::string value = line.substr(line.find("=") + 1);
    if (value.find("#") != std::string::npos) { value = value.substr(0, value.find("#")); }
    trim(key);
    trim(value);
    auto it = types_.find(key);
    if (it == types_.end()) {
        std::cerr << "Invalid key \"" + key + "\" and value \"" << value << "\"" << std::endl;
        return false;
    }
    bool ok = true;
    if (it->second == 'i' || it->second == 'f') {
        std::istringstream iss(value);
        double v;
        ok = static_cast<bool>(iss >> v);
    } else if (it->second == 'b') {
        ok = (value == "true" || value == "false" || value == "1" || value == "0");
        if (value == "1") { value = "true"; }
        if (value == "0") { value = "false"; }
    }
    if (!ok) {
        std::cerr << "Unsatisfiable value \"" + value + "\" for option \"" + key + "\"" << std::endl;
        return false;
    }
    values_[key] = value;
    return true;
}

int Config::getInt(const std::string& k) const { return std::atoi(values_.at(k).c_str()); }
float Config::getFloat(const std::string& k) const { return std::strtof(values_.at(k).c_str(), nullptr); }
bool Config::getBool(const std::string& k) const { return values_.at(k) == "true"; }
const std::string& Config::getString(const std::string& k) const { return values_.at(k); }

} // namespace mzhost
