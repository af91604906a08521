// Game record and `SelfPlay` line of the zero-server wire protocol, byte for byte as the reference writes them:
//   actor/actor_group.cpp:24-50      ThreadSharedData::outputGame ("SelfPlay <terminal> <data len> <game len> <return> <record> #")
//   actor/base_actor.cpp:42-66       BaseActor::getRecord / getActionInfo (tags EV, RE on resign, DLEN; P / V / R per move)
//   environment/base/base_env.h:207-233,303-313   loadFromEnvironment / toString / escapeSGFString (tag order = insertion order)
//   environment/go/go.h:129-133      KM tag; environment/base/base_env.h:363-367 SZ tag
//   actor/mcts.cpp:126-137           getSearchDistributionString ("action:count" of the visited root children, child order)
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include <zlib.h>

namespace mzhost {

// utils::compressString (utils/utils.h:34-47,49-57,84-87): gzip, then lower-case hex. The reference compresses through
// boost::iostreams::gzip_compressor (default level, 15-bit window); the deflate stream here is the same zlib call, the 10-byte gzip
// header is zlib's own (OS byte 3) where Boost writes its own (OS byte 255) — every gzip reader, the reference's included, accepts both.
inline std::string compressToHex(const std::string& in)
{
    if (in.empty()) { return in; }
    z_stream zs{};
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) { return std::string(); }
    std::string out(deflateBound(&zs, in.size()) + 32, '\0');
    zs.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(in.data())), zs.avail_in = static_cast<uInt>(in.size());
    zs.next_out = reinterpret_cast<Bytef*>(&out[0]), zs.avail_out = static_cast<uInt>(out.size());
    deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    static const char* digits = "0123456789abcdef";
    std::string hex(out.size() * 2, '0');
    for (size_t i = 0; i < out.size(); ++i) {
        const unsigned char c = static_cast<unsigned char>(out[i]);
        hex[2 * i] = digits[c >> 4], hex[2 * i + 1] = digits[c & 15];
    }
    return hex;
}

struct MoveRecord {
    int action = 0;
    int player = 1; // 1 = 'B', 2 = 'W'
    std::string policy, value, reward; // P, V, R
    bool cleared = false; // action info dropped after an intermediate sequence was sent (actor_group.cpp:40-47)
};

// zero_actor_intermediate_sequence_length > 0: long games are sent in pieces (actor_group.cpp:24-64,129-131)
struct SequenceConfig {
    int sequence_length = 0; // zero_actor_intermediate_sequence_length
    int unrolling_step = 5;  // learner_muzero_unrolling_step
    int n_step_return = 0;   // learner_n_step_return
};

// ThreadSharedData::calculateTrainingDataRange (actor_group.cpp:52-64)
inline std::pair<int, int> trainingDataRange(int game_length, bool env_terminal, const SequenceConfig& c)
{
    int data_start = 0, data_end = game_length - 1;
    if (c.sequence_length > 0) {
        data_end = std::max(0, (env_terminal ? data_end : data_end - c.unrolling_step - c.n_step_return));
        data_start = std::max(0, (env_terminal ? data_end - data_end % c.sequence_length : data_end + 1 - c.sequence_length));
        if (env_terminal && (data_end % c.sequence_length < c.unrolling_step + c.n_step_return)) { data_start = std::max(0, data_start - c.sequence_length); }
    }
    return {data_start, data_end};
}

// SlaveThread::handleSearchDone (actor_group.cpp:129-131): is an intermediate sequence due after a move that did not end the game?
inline bool intermediateSequenceDue(int game_length, const SequenceConfig& c)
{
    return c.sequence_length > 0 && game_length >= c.sequence_length && (game_length - c.n_step_return - c.unrolling_step) % c.sequence_length == 0;
}

inline char playerToChar(int p) { return p == 1 ? 'B' : (p == 2 ? 'W' : 'N'); } // environment/base/base_env.cpp:5-13

inline std::string escapeSGF(const std::string& str)
{
    static const std::string special = "()[]\\";
    std::string escaped;
    for (char c : str) {
        if (special.find(c) != std::string::npos) { escaped += '\\'; }
        escaped += c;
    }
    return escaped;
}

// MCTS::getSearchDistributionString: counts are floats streamed with operator<<
inline std::string searchDistribution(const int* actions, const float* counts, int num_children)
{
    std::ostringstream oss;
    bool first = true;
    for (int i = 0; i < num_children; ++i) {
        if (counts[i] == 0) { continue; }
        oss << (first ? "" : ",") << actions[i] << ":" << counts[i];
        first = false;
    }
    return oss.str();
}

// GumbelZero::getMCTSPolicy (actor/gumbel_zero.cpp:9-59) from the root child table: softmax of the completed-Q logits,
// entries below -38 dropped, printed in the iteration order of the same std::unordered_map<int, float> the reference fills
// (same libstdc++, same insertion sequence => same order). child_player = side to move.
// `normalized(i)` = MCTSNode::getNormalizedMean of root child i (mcts.cpp:40-53; the caller knows about rewards and value bounds).
template <class NormalizedMean>
inline std::string gumbelPolicy(const int* actions, const float* counts, const float* policy, const float* logit, const float* noise, int num_children, float root_value,
                                int child_player, int num_simulation, float sigma_visit_c, float sigma_scale_c, NormalizedMean normalized)
{
    float pi_sum = 0.0f, q_sum = 0.0f;
    for (int i = 0; i < num_children; ++i) {
        if (counts[i] == 0) { continue; }
        float value = normalized(i);
        pi_sum += policy[i];
        q_sum += policy[i] * value;
    }
    float value_pi = root_value;
    value_pi = (child_player == 2 ? -value_pi : value_pi);
    float non_visited_node_value = 1.0 / (1 + num_simulation) * (value_pi + (num_simulation / pi_sum) * q_sum);
    std::unordered_map<int, float> new_logits;
    float max_logit = -std::numeric_limits<float>::max();
    float max_child_count = 0;
    for (int i = 0; i < num_children; ++i) { max_child_count = fmax(max_child_count, counts[i]); }
    for (int i = 0; i < num_children; ++i) {
        float value = (counts[i] == 0 ? non_visited_node_value : normalized(i));
        float logit_without_noise = logit[i] - noise[i];
        float score = logit_without_noise + (sigma_visit_c + max_child_count) * sigma_scale_c * value;
        new_logits.insert({actions[i], score});
        max_logit = fmax(max_logit, score);
    }
    std::ostringstream oss;
    for (auto& l : new_logits) {
        l.second = l.second - max_logit;
        if (l.second < -38) { continue; }
        oss << (oss.str().empty() ? "" : ",") << l.first << ":" << exp(l.second);
    }
    return oss.str();
}

struct GameHeader {
    std::string game_name;  // Environment::name(): "go_9x9", "tictactoe"
    int board_size = 0;
    bool has_komi = false;
    float komi = 0.0f;
    std::string model_file; // config::nn_file_name
};

// what an Atari record carries beside the moves (AtariEnvLoader::loadFromEnvironment, atari.cpp:174-185; base_env.h:215-219)
struct AtariRecord {
    const std::vector<std::string>* observations = nullptr; // AtariEnv::observations_: one 3 x 96 x 96 byte screen per position, old ones emptied (atari.cpp:72-79)
    const std::vector<int>* lives_history = nullptr;        // lives before every action (+ the current ones)
    int seed = 0;                                           // SD tag
    float total_reward = 0.0f;                              // AtariEnv::getEvalScore
};

// eval_score: Environment::getEvalScore(false) of the final position; terminal: Environment::isTerminal().
// When the game is not terminal (resign) the side to move loses (base_actor.cpp:48-54, go.cpp:262-263).
inline std::string selfPlayLine(const GameHeader& h, const std::vector<MoveRecord>& moves, bool terminal, float eval_score, int turn_to_move,
                                const SequenceConfig& seq = SequenceConfig(), const AtariRecord* atari = nullptr, const float* unfinished_score = nullptr)
{
    // a game that is not over counts as resigned by the side to move (base_actor.cpp:48-54: getEvalScore(true)); Atari's score is the total reward either
    // way (atari.h:59); KillAllGo's getEvalScore ignores the resign flag and reads the position (killallgo.cpp:42-48): the caller passes it
    const float resign_score = (atari ? atari->total_reward : (unfinished_score ? *unfinished_score : (turn_to_move == 1 ? -1.0f : 1.0f)));
    if (atari) { eval_score = atari->total_reward; }
    std::vector<std::pair<std::string, std::string>> tags;
    tags.push_back({"GM", h.game_name});
    tags.push_back({"RE", std::to_string(eval_score)}); // loadFromEnvironment: std::to_string(env.getEvalScore())
    if (atari) {
        std::string all;
        for (const std::string& o : *atari->observations) { all += o; }
        tags.push_back({"OBS", compressToHex(all)});                  // base_env.h:215-219
        tags.push_back({"SD", std::to_string(atari->seed)});          // atari.cpp:177
    } else {
        tags.push_back({"OBS", ""});
        tags.push_back({"SZ", std::to_string(h.board_size)});
    }
    if (h.has_komi) { tags.push_back({"KM", std::to_string(h.komi)}); }
    tags.push_back({"EV", h.model_file.substr(h.model_file.find_last_of('/') + 1)});
    if (!terminal) {
        std::ostringstream oss;
        oss << resign_score;
        tags[1].second = oss.str();
    }
    const int game_length = static_cast<int>(moves.size());
    const std::pair<int, int> range = trainingDataRange(game_length, terminal, seq);
    tags.push_back({"DLEN", std::to_string(range.first) + "-" + std::to_string(range.second)});
    std::ostringstream rec;
    rec << "(;";
    for (const auto& t : tags) { rec << t.first << "[" << escapeSGF(t.second) << "]"; }
    for (size_t i = 0; i < moves.size(); ++i) {
        const MoveRecord& m = moves[i];
        rec << ";" << playerToChar(m.player) << "[" << m.action << "]";
        if (!m.cleared) { rec << "P[" << escapeSGF(m.policy) << "]V[" << escapeSGF(m.value) << "]R[" << escapeSGF(m.reward) << "]"; }
        // life lost since the position before (atari.cpp:178-184): added at record time, so it survives the clearing of sent action info
        if (atari && i > 0 && (*atari->lives_history)[i] < (*atari->lives_history)[i - 1]) { rec << "L[" << (*atari->lives_history)[i] << "]"; }
    }
    rec << ")";
    std::ostringstream oss;
    const bool is_terminal = (seq.sequence_length == 0 || terminal); // actor_group.cpp:30
    oss << "SelfPlay " << (is_terminal ? "true" : "false") << " " << (range.second - range.first + 1) << " " << game_length << " "
        << (terminal ? eval_score : resign_score) << " " << rec.str() << " #";
    return oss.str();
}

// after a non-terminal record was sent: "delete action info history if not complete record to save memory" (actor_group.cpp:40-47)
inline void clearSentActionInfo(std::vector<MoveRecord>& moves, bool terminal, const SequenceConfig& seq)
{
    if (seq.sequence_length == 0 || terminal) { return; }
    const std::pair<int, int> range = trainingDataRange(static_cast<int>(moves.size()), terminal, seq);
    for (int i = range.first; i <= range.second && i < static_cast<int>(moves.size()); ++i) { moves[i].cleared = true; }
}

} // namespace mzhost
