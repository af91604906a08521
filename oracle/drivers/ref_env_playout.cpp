// Differential-playout generator (the reference's own `-mode env_test` idea, console/mode_handler.cpp:167-192): plays random
// legal moves with the UNMODIFIED reference environment (compiled in place from /root/reference by oracle/Makefile) and dumps,
// before every move, the side to move, the legal action set, the feature planes under a rotation, and the move chosen; at the
// end of every game the terminal flag and the evaluation score. TEST INFRASTRUCTURE ONLY.
//
// usage: ref_env_playout <conf_str> <seed> <num_games> <max_moves_per_game> <out_file>
// record per step : i32 game, i32 step, i32 turn, i32 rotation, i32 action, i32 terminal_after, f32 score_after, u8 legal[A], u8 features[F]
#include "configuration.h"
#include "configure_loader.h"
#include "environment.h"
#include "rotation.h"
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>
#include <vector>

using namespace minizero;

int main(int argc, char** argv)
{
    if (argc < 6) {
        std::cerr << "usage: ref_env_playout <conf_str> <seed> <num_games> <max_moves_per_game> <out_file>" << std::endl;
        return 2;
    }
    env::setUpEnv();
    config::ConfigureLoader cl;
    config::setConfiguration(cl);
    if (!cl.loadFromString(argv[1])) { return 1; }
    std::mt19937 rng(static_cast<unsigned>(atoi(argv[2])));
    const int num_games = atoi(argv[3]), max_moves = atoi(argv[4]);
    FILE* f = fopen(argv[5], "wb");
    Environment env;
    const int A = env.getPolicySize();
    int F = 0;
    int first_action = 0;
    for (int g = 0; g < num_games; ++g) {
        env.reset();
        for (int step = 0; step < max_moves && !env.isTerminal(); ++step) {
            std::vector<uint8_t> legal(A, 0);
            std::vector<int> ids;
            for (int a = 0; a < A; ++a) {
                if (env.isLegalAction(Action(a, env.getTurn()))) {
                    legal[a] = 1;
                    ids.push_back(a);
                }
            }
            if (ids.empty()) { break; }
            const int rotation = static_cast<int>(rng() % 8);
            const std::vector<float> feats = env.getFeatures(static_cast<utils::Rotation>(rotation));
            F = static_cast<int>(feats.size());
            // mostly board moves: a uniformly random choice would pass far too early in Go
            int action = ids[rng() % ids.size()];
            if (ids.size() > 1 && action == A - 1 && (rng() % 8) != 0) { action = ids[rng() % (ids.size() - 1)]; }
            // Hex swap rule: the second move may repeat the first one (every other game refuses that, and then no coin is drawn)
            if (step == 1 && legal[first_action] && (rng() % 2) == 0) { action = first_action; }
            if (step == 0) { first_action = action; }
            const int turn = static_cast<int>(env.getTurn());
            if (!env.act(Action(action, env.getTurn()))) {
                std::cerr << "reference refused a move it reported legal" << std::endl;
                return 1;
            }
            const int32_t hdr[6] = {g, step, turn, rotation, action, env.isTerminal() ? 1 : 0};
            const float score = env.getEvalScore();
            fwrite(hdr, 4, 6, f);
            fwrite(&score, 4, 1, f);
            fwrite(legal.data(), 1, A, f);
            std::vector<uint8_t> fb(F);
            for (int k = 0; k < F; ++k) { fb[k] = (feats[k] != 0.0f); }
            fwrite(fb.data(), 1, F, f);
        }
    }
    fclose(f);
    std::cout << "A " << A << " F " << F << std::endl;
    return 0;
}
