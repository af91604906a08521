// mz_sp — drop-in for the reference's `-mode sp` executable (console/mode_handler.cpp:29-74,145-149):
//   mz_sp -conf_file <file> -conf_str "k=v:k=v" -mode sp
// plus `-mode record_test` (CPU only): formats a game described on stdin into a SelfPlay line, used by the tests to pin
// the wire format against lines recorded from the reference.
#include "config.h"
#include "record.h"
#include "worker.h"
#include <cstring>
#include <iostream>
#include <sstream>
#include <unistd.h>

namespace {

float hexFloat(const std::string& s)
{
    uint32_t bits = static_cast<uint32_t>(std::stoul(s, nullptr, 16));
    float f;
    std::memcpy(&f, &bits, 4);
    return f;
}

// stdin: "header <game_name> <board> <has_komi> <komi> <model_file> <terminal> <eval_hex> <turn_to_move> [<sequence_length> <unrolling_step> <n_step_return>]"
//        (with a sequence length, intermediate lines are printed after the moves where actor_group.cpp:129-131 sends them)
//        "move <player> <action> <root_mean_hex> <k> a:count_hex ..."   (one line per move), then EOF
int recordTest()
{
    mzhost::GameHeader h;
    std::vector<mzhost::MoveRecord> moves;
    bool terminal = true;
    float eval = 0.0f;
    int turn = 1;
    mzhost::SequenceConfig seq;
    auto maybe_intermediate = [&]() {
        if (!mzhost::intermediateSequenceDue(static_cast<int>(moves.size()), seq)) { return; }
        std::cout << mzhost::selfPlayLine(h, moves, false, 0.0f, moves.size() % 2 == 0 ? 1 : 2, seq) << std::endl;
        mzhost::clearSentActionInfo(moves, false, seq);
    };
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream iss(line);
        std::string kind;
        iss >> kind;
        if (kind == "header") {
            std::string eval_hex;
            int has_komi, term;
            iss >> h.game_name >> h.board_size >> has_komi >> h.komi >> h.model_file >> term >> eval_hex >> turn;
            h.has_komi = has_komi != 0, terminal = term != 0, eval = hexFloat(eval_hex);
            if (!(iss >> seq.sequence_length >> seq.unrolling_step >> seq.n_step_return)) { seq = mzhost::SequenceConfig(); }
        } else if (kind == "move") {
            mzhost::MoveRecord m;
            std::string mean_hex;
            int k;
            iss >> m.player >> m.action >> mean_hex >> k;
            std::vector<int> acts(k);
            std::vector<float> counts(k);
            for (int i = 0; i < k; ++i) {
                std::string tok;
                iss >> tok;
                acts[i] = std::stoi(tok.substr(0, tok.find(':')));
                counts[i] = hexFloat(tok.substr(tok.find(':') + 1));
            }
            m.policy = mzhost::searchDistribution(acts.data(), counts.data(), k);
            m.value = std::to_string(hexFloat(mean_hex));
            m.reward = "0";
            moves.push_back(m);
            maybe_intermediate();
        } else if (kind == "gmove") { // Gumbel: "gmove <player> <action> <root_mean_hex> <root_value_hex> <S> <visit_c> <scale_c> <k> a:count:mean:policy:logit:noise ..."
            mzhost::MoveRecord m;
            std::string mean_hex, value_hex;
            int k, sims;
            float visit_c, scale_c;
            iss >> m.player >> m.action >> mean_hex >> value_hex >> sims >> visit_c >> scale_c >> k;
            std::vector<int> acts(k);
            std::vector<float> f[5];
            for (auto& v : f) { v.resize(k); }
            for (int i = 0; i < k; ++i) {
                std::string tok, part;
                iss >> tok;
                std::istringstream ts(tok);
                std::getline(ts, part, ':');
                acts[i] = std::stoi(part);
                for (auto& v : f) {
                    std::getline(ts, part, ':');
                    v[i] = hexFloat(part);
                }
            }
            const int player = m.player;
            m.policy = mzhost::gumbelPolicy(acts.data(), f[0].data(), f[2].data(), f[3].data(), f[4].data(), k, hexFloat(value_hex), m.player, sims, visit_c, scale_c, [&](int i) {
                float value = 0.0f + 1.0f * f[1][i]; // board games: reward 0, discount 1, no rescale (mcts.cpp:40-53)
                value = (player == 2 ? -value : value);
                return (value * f[0][i] - 0.0f) / (f[0][i] + 0.0f);
            });
            m.value = std::to_string(hexFloat(mean_hex));
            m.reward = "0";
            moves.push_back(m);
        }
    }
    std::cout << mzhost::selfPlayLine(h, moves, terminal, eval, turn, seq) << std::endl;
    return 0;
}

} // namespace

int main(int argc, char** argv)
{
    std::string mode, conf_file, conf_str;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string flag = argv[i];
        if (flag == "-mode") {
            mode = argv[i + 1];
        } else if (flag == "-conf_file") {
            conf_file = argv[i + 1];
        } else if (flag == "-conf_str") {
            conf_str = argv[i + 1];
        } else {
            std::cerr << "unknown flag " << flag << std::endl;
            return -1;
        }
    }
    if (mode == "record_test") { return recordTest(); }
    if (mode != "sp" && mode != "rng_test") {
        std::cerr << "usage: mz_sp -mode sp [-conf_file F] [-conf_str \"k=v:...\"]   (other modes of the reference binary are out of scope)" << std::endl;
        return -1;
    }
    mzhost::Config cfg;
    if (!conf_file.empty() && !cfg.loadFromFile(conf_file)) {
        std::cerr << "Failed to load configuration file." << std::endl;
        return -1;
    }
    if (!conf_str.empty() && !cfg.loadFromString(conf_str)) {
        std::cerr << "Failed to load configuration string." << std::endl;
        return -1;
    }
    if (cfg.getString("nn_type_name") != "alphazero" && cfg.getString("nn_type_name") != "muzero") {
        std::cerr << "this worker implements the alphazero and (board-game) muzero self-play paths" << std::endl;
        return -1;
    }
    if (mode == "rng_test") { // CPU only: prints the host's draw sequence for root tables given on stdin
        mzhost::Worker worker(cfg, 1);
        return worker.rngTest(std::cin);
    }
    // stdout belongs to the wire protocol: the server drops the connection on anything but `SelfPlay` lines
    // (zero/zero_server.cpp:130-139). Libraries in this process (NCCL prints its version banner to stdout) must not be able
    // to write there, so the real stdout is kept on a private descriptor and fd 1 is pointed at stderr.
    const int wire_fd = dup(1);
    if (wire_fd < 0 || dup2(2, 1) < 0) {
        std::cerr << "cannot set up the output descriptors" << std::endl;
        return -1;
    }
    int rc;
    {
        mzhost::Worker worker(cfg, wire_fd);
        rc = worker.run();
    } // engines destroyed here
    // the stdin reader thread may still sit inside std::getline holding std::cin's lock: leave without running the iostream
    // destructors, which would wait for it (the reference's `quit` ends the process with exit(0), actor_group.cpp:250)
    std::cerr.flush();
    _exit(rc);
}
