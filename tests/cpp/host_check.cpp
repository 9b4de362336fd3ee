// Test driver for include/qwen3_transformer.hpp (the C++ host mirror of the reference interface).
//   host_check sample <vocab> <temperature> <topp> <seed> <n> <logits.f32>   -> n sampled ids, then the RNG state
//   host_check errors <missing.bin> <valid.bin>                               -> error codes of the two constructions
//   host_check generate <ckpt.bin> <max_new> <tok> [tok ...]                  -> greedy generate() + decode_greedy() ids (needs a GPU)
//   host_check chat <use_prefill> <temperature> <topp> <seed> <max_new>        -> replies of a 3-turn chat on a fake transformer, RNG state, #forwards
//   host_check tokenize <ckpt path> <vocab> <thinking> <texts file>           -> per line of the file: its token ids; then a summary line
//   host_check render <ckpt path> <vocab> <thinking> <pos> <system|-> <user>  -> the rendered prompt
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "qwen3_transformer.hpp"

// deterministic stand-in for the transformer (host-logic checks only): logits are an integer hash of (i, token, pos)
struct FakeTransformer {
    struct Cfg {
        int seq_len, vocab_size;
    } cfg{40, 64};
    std::vector<float> logits = std::vector<float>(64);
    size_t forwards = 0, prefills = 0;
    const Cfg &get_config() const { return cfg; }
    const std::vector<float> &forward(size_t token, size_t pos) {
        forwards++;
        for (uint32_t i = 0; i < 64; i++)
            logits[i] = (float)(((i * 2654435761u + (uint32_t)token * 40503u + (uint32_t)pos * 69069u) >> 8) & 1023u) / 64.0f;
        return logits;
    }
    const std::vector<float> &prefill(const std::vector<int> &tokens, size_t pos0) { // q3_prefill's contract
        prefills++;
        const size_t before = forwards;
        for (size_t i = 0; i < tokens.size(); i++) forward((size_t)tokens[i], pos0 + i);
        forwards = before;
        return logits;
    }
};

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    const std::string mode = argv[1];
    if (mode == "sample" && argc == 8) {
        const size_t V = std::strtoul(argv[2], nullptr, 10);
        qwen3::Sampler s(V, std::strtof(argv[3], nullptr), std::strtof(argv[4], nullptr), std::strtoull(argv[5], nullptr, 10));
        const size_t n = std::strtoul(argv[6], nullptr, 10);
        FILE *f = std::fopen(argv[7], "rb");
        if (!f) return 3;
        std::vector<float> logits(V);
        for (size_t i = 0; i < n; i++) {
            if (std::fread(logits.data(), 4, V, f) != V) return 4;
            std::printf("%zu\n", s.sample(logits));
        }
        std::fclose(f);
        std::printf("%llu\n", (unsigned long long)s.rng_state);
        return 0;
    }
    if (mode == "errors" && argc == 4) {
        for (int i = 2; i < 4; i++) {
            try {
                qwen3::Transformer t = qwen3::TransformerBuilder(argv[i]).with_ctx_length(std::nullopt).build();
                std::printf("0 ok vocab=%d\n", t.get_config().vocab_size);
            } catch (const qwen3::Error &e) {
                std::printf("%d %s\n", e.code(), e.what());
            }
        }
        return 0;
    }
    if (mode == "generate" && argc >= 5) {
        const size_t max_new = std::strtoul(argv[3], nullptr, 10);
        std::vector<size_t> prompt;
        for (int i = 4; i < argc; i++) prompt.push_back(std::strtoul(argv[i], nullptr, 10));
        qwen3::Transformer t = qwen3::TransformerBuilder(argv[2]).with_ctx_length(64).build();
        qwen3::Sampler s((size_t)t.get_config().vocab_size, 0.0f, 0.9f, 0);
        for (size_t tok : qwen3::generate(t, s, prompt, max_new)) std::printf("%zu ", tok);
        std::printf("\n");
        t.reset();
        for (int tok : t.decode_greedy(prompt.back(), prompt.size() - 1, max_new)) std::printf("%d ", tok);
        std::printf("\n");
        try {
            t.forward((size_t)t.get_config().vocab_size, 0); // out of range: the reference panics
            std::printf("no error\n");
        } catch (const qwen3::Error &e) {
            std::printf("%d\n", e.code());
        }
        return 0;
    }
    if (mode == "chat" && argc == 7) {
        FakeTransformer t;
        qwen3::Sampler s(64, std::strtof(argv[3], nullptr), std::strtof(argv[4], nullptr), std::strtoull(argv[5], nullptr, 10));
        const std::vector<std::vector<size_t>> turns = {{5, 9, 20, 31}, {7, 7, 30}, {11}};
        for (const auto &reply : qwen3::chat(t, s, turns, -1, -1, std::atoi(argv[2]) != 0, std::strtoul(argv[6], nullptr, 10))) {
            for (size_t tok : reply) std::printf("%zu ", tok);
            std::printf("\n");
        }
        std::printf("rng %llu forwards %zu prefills %zu\n", (unsigned long long)s.rng_state, t.forwards, t.prefills);
        return 0;
    }
    if (mode == "tokenize" && argc == 6) {
        qwen3::Tokenizer tok(argv[2], std::strtoul(argv[3], nullptr, 10), std::atoi(argv[4]) != 0);
        FILE *f = std::fopen(argv[5], "rb");
        if (!f) return 3;
        std::string all;
        char buf[4096];
        size_t n;
        while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) all.append(buf, n);
        std::fclose(f);
        size_t start = 0;
        while (start <= all.size()) { // one text per line (texts contain no newline)
            size_t nl = all.find('\n', start);
            if (nl == std::string::npos) nl = all.size();
            std::string line = all.substr(start, nl - start);
            std::string back;
            for (size_t t : tok.encode(line)) {
                std::printf("%zu ", t);
                back += tok.decode(t);
            }
            std::printf("| %d\n", back == line ? 1 : 0);
            if (nl == all.size()) break;
            start = nl + 1;
        }
        std::printf("meta %u %u %u %zu\n", tok.max_token_length, tok.bos_token_id, tok.eos_token_id, tok.vocab.size());
        return 0;
    }
    if (mode == "render" && argc == 8) {
        qwen3::Tokenizer tok(argv[2], std::strtoul(argv[3], nullptr, 10), std::atoi(argv[4]) != 0);
        std::optional<std::string> sys;
        if (std::string(argv[6]) != "-") sys = argv[6];
        std::fputs(qwen3::render_prompt(std::strtoul(argv[5], nullptr, 10), sys, argv[7], tok).c_str(), stdout);
        return 0;
    }
    return 2;
}
