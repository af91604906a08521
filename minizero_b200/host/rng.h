// Host-side random draws of the self-play path, made with the same libstdc++ engine and distributions, in the same
// order, as the reference (utils/random.h:9-41, utils/random.cpp:5-7): exact reproduction of a
// seed means calling the same objects in the same sequence, so all randomness of a search is drawn here and shipped
// to the device (rotations, root noise) or consumed on the host (move choice, resign switch).
#pragma once
#include <cmath>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

namespace mzhost {

class Random {
public:
    void seed(int s) { generator_.seed(s); }
    int randInt() { return int_distribution_(generator_); }
    double randReal(double range = 1.0f) { return real_distribution_(generator_) * range; }

    // utils/random.h:15-24
    std::vector<float> randDirichlet(float alpha, int size)
    {
        std::vector<float> dirichlet;
        std::gamma_distribution<float> gamma_distribution(alpha);
        for (int i = 0; i < size; ++i) { dirichlet.emplace_back(gamma_distribution(generator_)); }
        float sum = std::accumulate(dirichlet.begin(), dirichlet.end(), 0.0f);
        if (sum < std::numeric_limits<float>::min()) { return dirichlet; }
        for (int i = 0; i < size; ++i) { dirichlet[i] /= sum; }
        return dirichlet;
    }

    // utils/random.h:26-36
    std::vector<float> randGumbel(int size)
    {
        std::extreme_value_distribution<float> gumbel_distribution(0.0, 1.0);
        std::vector<float> gumbel;
        for (int i = 0; i < size; ++i) {
            float value = gumbel_distribution(generator_);
            while (std::isinf(value)) { value = gumbel_distribution(generator_); }
            gumbel.emplace_back(value);
        }
        return gumbel;
    }

private:
    std::mt19937 generator_;
    std::uniform_int_distribution<int> int_distribution_;
    std::uniform_real_distribution<double> real_distribution_;
};

} // namespace mzhost
