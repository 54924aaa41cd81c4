#include "corpus.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <string_view>
#include <unordered_map>

#include "log.h"

namespace adapm {

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  ADAPM_CHECK(f.good(), "cannot open corpus file '" << path << "'");
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0, std::ios::beg);
  std::string s((size_t)n, '\0');
  f.read(&s[0], n);
  return s;
}

// calls fn(line_no, begin, end) for every line (without the terminator)
template <class F>
void for_each_line(const std::string& buf, F&& fn) {
  size_t b = 0;
  int64_t li = 0;
  const size_t n = buf.size();
  while (b < n) {
    size_t e = buf.find('\n', b);
    if (e == std::string::npos) e = n;
    fn(li++, b, e);
    b = e + 1;
  }
}

template <class F>
void for_each_word(const std::string& buf, size_t b, size_t e, F&& fn) {
  size_t i = b;
  while (i < e) {
    while (i < e && is_space(buf[i])) ++i;
    size_t j = i;
    while (j < e && !is_space(buf[j])) ++j;
    if (j > i) fn(std::string_view(buf.data() + i, j - i));
    i = j;
  }
}

// splitmix64 / xorshift: small, fast, reproducible across platforms
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed ? seed : 0x9E3779B97F4A7C15ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

}  // namespace

struct Corpus::Index {
  std::unordered_map<std::string, int32_t> map;
};

std::shared_ptr<Corpus> Corpus::build(const std::string& path, int64_t min_count) {
  const std::string buf = read_file(path);
  struct Ent { int64_t count, first; };
  std::unordered_map<std::string_view, Ent> cnt;
  cnt.reserve(1 << 16);
  int64_t lines = 0, order = 0;
  for_each_line(buf, [&](int64_t, size_t b, size_t e) {
    ++lines;
    for_each_word(buf, b, e, [&](std::string_view w) {
      auto it = cnt.find(w);
      if (it == cnt.end()) cnt.emplace(w, Ent{1, order++});
      else ++it->second.count;
    });
  });
  std::vector<std::pair<std::string_view, Ent>> items;
  items.reserve(cnt.size());
  for (auto& p : cnt)
    if (p.second.count >= min_count) items.push_back(p);
  std::sort(items.begin(), items.end(), [](const auto& a, const auto& b) {
    if (a.second.count != b.second.count) return a.second.count > b.second.count;
    return a.second.first < b.second.first;   // ties: first occurrence first (a stable sort of the insertion order)
  });
  auto c = std::make_shared<Corpus>();
  c->words_.reserve(items.size() + 1);
  c->counts_.reserve(items.size() + 1);
  c->words_.push_back("</s>");
  c->counts_.push_back(std::max<int64_t>(1, lines));
  for (auto& p : items) {
    c->words_.emplace_back(p.first);
    c->counts_.push_back(p.second.count);
  }
  c->index_words();
  return c;
}

std::shared_ptr<Corpus> Corpus::from_vocab(std::vector<std::string> words, std::vector<int64_t> counts) {
  ADAPM_CHECK(words.size() == counts.size() && !words.empty(), "vocabulary: words and counts must have the same length");
  auto c = std::make_shared<Corpus>();
  c->words_ = std::move(words);
  c->counts_ = std::move(counts);
  c->index_words();
  return c;
}

void Corpus::index_words() {
  index_ = std::make_shared<Index>();
  index_->map.reserve(words_.size() * 2);
  for (size_t i = 0; i < words_.size(); ++i) index_->map.emplace(words_[i], (int32_t)i);
}

int64_t Corpus::lookup(const std::string& w) const {
  auto it = index_->map.find(w);
  return it == index_->map.end() ? -1 : it->second;
}

void Corpus::encode(const std::string& path, int rank, int world) {
  ADAPM_CHECK(world >= 1 && rank >= 0 && rank < world, "corpus: bad rank/world");
  const std::string buf = read_file(path);
  tokens_.clear();
  sent_off_.assign(1, 0);
  std::string tmp;
  for_each_line(buf, [&](int64_t li, size_t b, size_t e) {
    if (li % world != rank) return;
    const size_t before = tokens_.size();
    for_each_word(buf, b, e, [&](std::string_view w) {
      tmp.assign(w.data(), w.size());
      auto it = index_->map.find(tmp);
      if (it != index_->map.end()) tokens_.push_back(it->second);
    });
    if (tokens_.size() > before) sent_off_.push_back((int64_t)tokens_.size());
  });
}

// ------------------------------------------------------------------------------------- pair stream
PairStream::PairStream(std::shared_ptr<Corpus> corpus, int window, double subsample, int64_t batch_pairs, uint64_t seed,
                       int queue_depth)
    : corpus_(std::move(corpus)), window_(window), subsample_(subsample), batch_(batch_pairs), seed_(seed) {
  ADAPM_CHECK(window_ >= 1 && batch_ >= 1 && queue_depth >= 1, "PairStream: bad window / batch size / queue depth");
  ring_.resize((size_t)queue_depth);
  for (auto& s : ring_) s.keys.resize((size_t)(2 * batch_));
}

PairStream::~PairStream() { stop(); }

void PairStream::stop() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    abort_ = true;
  }
  cv_put_.notify_all();
  cv_get_.notify_all();
  if (thread_.joinable()) thread_.join();
  abort_ = false;
}

void PairStream::start_epoch(uint64_t epoch) {
  stop();
  {
    std::lock_guard<std::mutex> lk(mu_);
    head_ = tail_ = count_ = 0;
    done_ = false;
  }
  thread_ = std::thread([this, epoch] { run(epoch); });
}

void PairStream::run(uint64_t epoch) {
  Rng rng(seed_ * 0x2545F4914F6CDD1Dull + epoch * 0x9E3779B97F4A7C15ull + 1);
  const int32_t* tok = corpus_->tokens();
  const int64_t* off = corpus_->sentence_offsets();
  const int64_t ns = corpus_->num_sentences();
  const std::vector<int64_t>& counts = corpus_->counts();
  double total = 0;
  for (int64_t c : counts) total += (double)c;
  // keep probability per word (word2vec.c rule): (sqrt(f / t) + 1) * t / f
  std::vector<float> keep;
  if (subsample_ > 0) {
    keep.resize(counts.size());
    for (size_t i = 0; i < counts.size(); ++i) {
      const double fr = (double)counts[i] / total;
      keep[i] = (float)((std::sqrt(fr / subsample_) + 1.0) * subsample_ / fr);
    }
  }
  std::vector<Key> cur((size_t)(2 * batch_)), uniq;
  uniq.reserve((size_t)(2 * batch_));
  int64_t n = 0;
  std::vector<int32_t> s;

  auto publish = [&](int64_t valid) -> bool {
    if (valid < batch_) {   // pad a short last batch by repeating its pairs (fixed-size batches for the fused kernel)
      for (int64_t i = valid; i < batch_; ++i) {
        cur[(size_t)i] = cur[(size_t)(i % valid)];
        cur[(size_t)(batch_ + i)] = cur[(size_t)(batch_ + i % valid)];
      }
    }
    // distinct keys of the batch (before taking the lock: this is the expensive part)
    if (seen_.empty()) seen_.assign((size_t)(2 * corpus_->vocab_size()), 0u);
    if (++seen_epoch_ == 0) { std::fill(seen_.begin(), seen_.end(), 0u); seen_epoch_ = 1; }
    uniq.clear();
    for (int64_t i = 0; i < 2 * batch_; ++i) {
      const Key k = cur[(size_t)i];
      if (seen_[(size_t)k] != seen_epoch_) { seen_[(size_t)k] = seen_epoch_; uniq.push_back(k); }
    }
    std::unique_lock<std::mutex> lk(mu_);
    cv_put_.wait(lk, [&] { return count_ < ring_.size() || abort_; });
    if (abort_) return false;
    Slot& sl = ring_[tail_];
    sl.keys.swap(cur);
    sl.uniq.swap(uniq);
    sl.valid = valid;
    tail_ = (tail_ + 1) % ring_.size();
    ++count_;
    produced_.fetch_add((uint64_t)valid);
    lk.unlock();
    cv_get_.notify_one();
    return true;
  };

  bool ok = true;
  for (int64_t si = 0; si < ns && ok; ++si) {
    s.clear();
    for (int64_t t = off[si]; t < off[si + 1]; ++t) {
      const int32_t w = tok[t];
      if (subsample_ > 0 && keep[(size_t)w] < 1.f && rng.uniform() >= keep[(size_t)w]) continue;
      s.push_back(w);
    }
    const int64_t L = (int64_t)s.size();
    if (L < 2) continue;
    for (int64_t pos = 0; pos < L && ok; ++pos) {
      const int64_t b = (int64_t)(rng.next() % (uint64_t)window_);   // window shrink in [0, window)
      const int64_t r = window_ - b;
      const int64_t lo = std::max<int64_t>(0, pos - r), hi = std::min<int64_t>(L, pos + r + 1);
      const Key target = 2 * (Key)s[(size_t)pos] + 1;               // syn1 key of the centre word
      for (int64_t j = lo; j < hi; ++j) {
        if (j == pos) continue;
        cur[(size_t)n] = 2 * (Key)s[(size_t)j];                       // syn0 key of the context word
        cur[(size_t)(batch_ + n)] = target;
        if (++n == batch_) {
          ok = publish(batch_);
          n = 0;
          if (!ok) break;
        }
      }
    }
  }
  if (ok && n > 0) publish(n);
  {
    std::lock_guard<std::mutex> lk(mu_);
    done_ = true;
  }
  cv_get_.notify_all();
}

int64_t PairStream::next(Key* out) { return next_with_unique(out, nullptr, nullptr); }

int64_t PairStream::next_with_unique(Key* out, Key* uniq, int64_t* n_uniq) {
  std::unique_lock<std::mutex> lk(mu_);
  cv_get_.wait(lk, [&] { return count_ > 0 || done_ || abort_; });
  if (count_ == 0) return 0;
  Slot& sl = ring_[head_];
  const int64_t valid = sl.valid;
  lk.unlock();
  memcpy(out, sl.keys.data(), (size_t)(2 * batch_) * sizeof(Key));   // the slot stays reserved while we copy
  if (uniq) {
    memcpy(uniq, sl.uniq.data(), sl.uniq.size() * sizeof(Key));
    *n_uniq = (int64_t)sl.uniq.size();
  }
  lk.lock();
  head_ = (head_ + 1) % ring_.size();
  --count_;
  lk.unlock();
  cv_put_.notify_one();
  return valid;
}

}  // namespace adapm
