#include "fabric.h"

#include <dirent.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>

namespace adapm {

namespace { bool process_dead(int pid); }

namespace {

void init_control(ControlBlock* c, int world, int workers) {
  memset((void*)c, 0, sizeof(ControlBlock));
  c->world = world;
  c->workers = workers;
  c->sync_barrier.init();
  c->node_barrier.init();
  c->worker_barrier.init();
  for (int r = 0; r < MAX_RANKS; ++r)
    for (int w = 0; w < MAX_LOCAL_WORKERS; ++w) c->ranks[r].worker_clock[w].store(0);
  c->magic = kControlMagic;
  c->initialized.store(1, std::memory_order_release);
}

// ------------------------------------------------------------------ in-process world
struct InProcWorld {
  std::mutex mu;
  ControlBlock* ctl = nullptr;
  std::vector<char*> heaps;
  std::vector<uint64_t> bytes;
  std::vector<bool> is_cuda;
  int refs = 0;
  int world = 0;
};
std::mutex g_mu;
std::map<std::string, std::shared_ptr<InProcWorld>> g_worlds;

class InProcFabric : public Fabric {
 public:
  InProcFabric(const Options& opt) {
    rank_ = opt.rank; world_ = opt.world; cuda_ = opt.backend == "cuda";
    timeout_s_ = opt.wait_timeout_s;
    device_ = opt.device;
    job_ = opt.job;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_worlds.find(job_);
    if (it == g_worlds.end()) {
      w_ = std::make_shared<InProcWorld>();
      w_->ctl = (ControlBlock*)aligned_alloc(64, (sizeof(ControlBlock) + 63) / 64 * 64);
      init_control(w_->ctl, world_, opt.workers);
      w_->heaps.assign(world_, nullptr);
      w_->bytes.assign(world_, 0);
      w_->is_cuda.assign(world_, false);
      w_->world = world_;
      g_worlds[job_] = w_;
    } else {
      w_ = it->second;
      ADAPM_CHECK(w_->world == world_, "inproc world size mismatch for job " << job_);
    }
    w_->refs++;
    ctl_ = w_->ctl;
    heaps_.assign(world_, nullptr);
    if (cuda_) {
      ADAPM_CHECK(cudamem::available(), "backend=cuda but no CUDA device is visible");
      if (device_ < 0) device_ = 0;
      cudamem::set_device(device_);
    }
  }
  ~InProcFabric() override {
    if (my_heap_) {
      if (cuda_) { cudamem::set_device(device_); cudamem::free_dev(my_heap_); }
      else free(my_heap_);
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (--w_->refs == 0) {
      free(w_->ctl);
      g_worlds.erase(job_);
    }
  }
  void allocate_heaps(uint64_t bytes) override {
    if (cuda_) {
      cudamem::set_device(device_);
      my_heap_ = cudamem::alloc_zeroed(bytes);
      ctl_->ranks[rank_].device = device_;
    } else {
      uint64_t b = (bytes + 4095) / 4096 * 4096;
      my_heap_ = (char*)aligned_alloc(4096, b);
      ADAPM_CHECK(my_heap_, "out of memory allocating " << b << " bytes");
      memset(my_heap_, 0, b);
    }
    {
      std::lock_guard<std::mutex> lk(w_->mu);
      w_->heaps[rank_] = my_heap_;
      w_->bytes[rank_] = bytes;
    }
    node_barrier("inproc heap exchange");
    {
      std::lock_guard<std::mutex> lk(w_->mu);
      for (int r = 0; r < world_; ++r) heaps_[r] = w_->heaps[r];
    }
    if (cuda_) {
      for (int r = 0; r < world_; ++r) {
        int pd = ctl_->ranks[r].device;
        if (r != rank_ && pd != device_) cudamem::enable_peer(device_, pd);
      }
    }
    node_barrier("inproc heap mapped");
  }

 private:
  std::string job_;
  std::shared_ptr<InProcWorld> w_;
  char* my_heap_ = nullptr;
};

// ------------------------------------------------------------------ POSIX shm world
// Crashed jobs cannot unlink their POSIX-shm segments. Rank 0 of every new job removes the segments of jobs whose
// processes are all gone (and that are older than two minutes, so that a job that is just starting is never touched).
void gc_stale_shm_segments() {
  DIR* d = opendir("/dev/shm");
  if (!d) return;
  std::vector<std::string> ctl;
  while (dirent* e = readdir(d)) {
    const std::string n = e->d_name;
    if (n.size() > 10 && n.compare(0, 6, "adapm_") == 0 && n.compare(n.size() - 4, 4, "_ctl") == 0) ctl.push_back(n);
  }
  closedir(d);
  const time_t now = time(nullptr);
  long min_age = 120;
  if (const char* e = getenv("ADAPM_SHM_GC_AGE_S")) min_age = atol(e);
  for (const std::string& n : ctl) {
    const std::string path = "/dev/shm/" + n;
    struct stat st;
    if (stat(path.c_str(), &st) != 0 || now - st.st_ctime < min_age || (size_t)st.st_size < sizeof(ControlBlock)) continue;
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) continue;
    void* p = mmap(nullptr, sizeof(ControlBlock), PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) continue;
    const ControlBlock* c = (const ControlBlock*)p;
    bool stale = c->magic == kControlMagic && c->world >= 1 && c->world <= MAX_RANKS;
    int world = stale ? c->world : 0;
    for (int r = 0; r < world && stale; ++r) {
      const int pid = c->ranks[r].pid;
      if (pid > 0 && !process_dead(pid)) stale = false;
    }
    munmap(p, sizeof(ControlBlock));
    if (!stale) continue;
    const std::string prefix = "/" + n.substr(0, n.size() - 4);
    for (int r = 0; r < world; ++r) shm_unlink((prefix + "_h" + std::to_string(r)).c_str());
    shm_unlink((prefix + "_ctl").c_str());
    VLOG(1, "removed the shared-memory segments of the dead job " << prefix.substr(7));
  }
}

class ShmFabric : public Fabric {
 public:
  ShmFabric(const Options& opt) {
    rank_ = opt.rank; world_ = opt.world; cuda_ = opt.backend == "cuda";
    timeout_s_ = opt.wait_timeout_s;
    device_ = opt.device;
    prefix_ = "/adapm_" + opt.job;
    heaps_.assign(world_, nullptr);
    heap_bytes_.assign(world_, 0);
    const std::string name = prefix_ + "_ctl";
    const size_t sz = (sizeof(ControlBlock) + 4095) / 4096 * 4096;
    if (rank_ == 0) {
      gc_stale_shm_segments();
      shm_unlink(name.c_str());
      int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      ADAPM_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << strerror(errno));
      ADAPM_CHECK(ftruncate(fd, sz) == 0, "ftruncate failed");
      void* p = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      ADAPM_CHECK(p != MAP_FAILED, "mmap control failed");
      ctl_ = (ControlBlock*)p;
      init_control(ctl_, world_, opt.workers);
    } else {
      auto t0 = std::chrono::steady_clock::now();
      for (;;) {
        int fd = shm_open(name.c_str(), O_RDWR, 0600);
        if (fd >= 0) {
          struct stat st;
          if (fstat(fd, &st) == 0 && (size_t)st.st_size >= sz) {
            void* p = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            close(fd);
            ADAPM_CHECK(p != MAP_FAILED, "mmap control failed");
            ControlBlock* c = (ControlBlock*)p;
            if (c->initialized.load(std::memory_order_acquire) == 1 && c->magic == kControlMagic) {
              ctl_ = c;
              break;
            }
            munmap(p, sz);
          } else {
            close(fd);
          }
        }
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ADAPM_CHECK(el < timeout_s_, "timed out waiting for rank 0's control block " << name);
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
      }
      ADAPM_CHECK(ctl_->world == world_, "world size mismatch with rank 0");
    }
    ctl_sz_ = sz;
    peers_are_processes_ = true;
    ctl_->ranks[rank_].pid = (int)getpid();
    ctl_->ranks[rank_].attached.store(1);
    if (cuda_) {
      ADAPM_CHECK(cudamem::available(), "backend=cuda but no CUDA device is visible");
      if (device_ < 0) device_ = rank_ % cudamem::device_count();
      cudamem::set_device(device_);
    }
    // everybody attached? (also detects ranks that mapped a stale segment)
    node_barrier("shm attach");
  }

  ~ShmFabric() override {
    stop_failure_detector();
    for (int r = 0; r < world_; ++r) {
      if (!heaps_[r]) continue;
      if (cuda_) {
        if (r == rank_) cudamem::free_dev(heaps_[r]);
        else cudamem::close_handle(heaps_[r]);
      } else {
        munmap(heaps_[r], heap_bytes_[r]);
      }
    }
    if (!cuda_ && my_heap_created_) shm_unlink((prefix_ + "_h" + std::to_string(rank_)).c_str());
    if (ctl_) munmap((void*)ctl_, ctl_sz_);
    if (rank_ == 0) shm_unlink((prefix_ + "_ctl").c_str());
  }

  void allocate_heaps(uint64_t bytes) override {
    const uint64_t b = (bytes + 4095) / 4096 * 4096;
    RankControl& me = ctl_->ranks[rank_];
    if (cuda_) {
      cudamem::set_device(device_);
      heaps_[rank_] = cudamem::alloc_zeroed(b);
      cudamem::export_handle(heaps_[rank_], me.ipc_handle);
      me.device = device_;
    } else {
      const std::string name = prefix_ + "_h" + std::to_string(rank_);
      shm_unlink(name.c_str());
      int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
      ADAPM_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << strerror(errno));
      ADAPM_CHECK(ftruncate(fd, b) == 0, "ftruncate(" << b << ") failed: " << strerror(errno));
      void* p = mmap(nullptr, b, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      close(fd);
      ADAPM_CHECK(p != MAP_FAILED, "mmap heap failed");
      heaps_[rank_] = (char*)p;  // fresh shm pages are zero
      my_heap_created_ = true;
    }
    me.heap_bytes = b;
    heap_bytes_[rank_] = b;
    node_barrier("shm heap exchange");
    for (int r = 0; r < world_; ++r) {
      if (r == rank_) continue;
      RankControl& pr = ctl_->ranks[r];
      heap_bytes_[r] = pr.heap_bytes;
      if (cuda_) {
        if (pr.device != device_) cudamem::enable_peer(device_, pr.device);
        heaps_[r] = cudamem::import_handle(pr.ipc_handle);
      } else {
        const std::string name = prefix_ + "_h" + std::to_string(r);
        int fd = shm_open(name.c_str(), O_RDWR, 0600);
        ADAPM_CHECK(fd >= 0, "shm_open(" << name << ") failed: " << strerror(errno));
        void* p = mmap(nullptr, pr.heap_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        ADAPM_CHECK(p != MAP_FAILED, "mmap peer heap failed");
        heaps_[r] = (char*)p;
      }
    }
    node_barrier("shm heap mapped");
  }

 private:
  std::string prefix_;
  size_t ctl_sz_ = 0;
  std::vector<uint64_t> heap_bytes_;
  bool my_heap_created_ = false;
};

}  // namespace

std::shared_ptr<Fabric> Fabric::create(const Options& opt) {
  ADAPM_CHECK(opt.world >= 1 && opt.world <= MAX_RANKS, "world size must be in [1," << MAX_RANKS << "]");
  ADAPM_CHECK(opt.rank >= 0 && opt.rank < opt.world, "bad rank " << opt.rank);
  ADAPM_CHECK(opt.workers >= 1 && opt.workers <= MAX_LOCAL_WORKERS, "workers per rank must be in [1," << MAX_LOCAL_WORKERS << "]");
  if (opt.fabric == "inproc") return std::make_shared<InProcFabric>(opt);
  if (opt.fabric == "shm") return std::make_shared<ShmFabric>(opt);
  throw Error("unknown fabric '" + opt.fabric + "' (inproc|shm)");
}

// ------------------------------------------------------------------ failure detector
namespace {
bool process_dead(int pid) {
  if (pid <= 0) return false;
  if (kill(pid, 0) != 0 && errno == ESRCH) return true;
  // a process that exited but was not reaped yet (zombie) still answers kill(pid, 0)
  char path[64];
  snprintf(path, sizeof(path), "/proc/%d/stat", pid);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char buf[512];
  size_t n = fread(buf, 1, sizeof(buf) - 1, f);
  fclose(f);
  buf[n] = 0;
  const char* p = strrchr(buf, ')');   // the command name may contain spaces / parentheses
  return p && p[1] == ' ' && (p[2] == 'Z' || p[2] == 'X');
}
}  // namespace

void Fabric::start_failure_detector(int period_ms) {
  if (!peers_are_processes_ || world_ <= 1 || fd_thread_.joinable()) return;
  fd_stop_.store(false);
  fd_thread_ = std::thread([this, period_ms] {
    while (!fd_stop_.load(std::memory_order_acquire)) {
      for (int r = 0; r < world_ && dead_peer_.load() < 0; ++r) {
        if (r == rank_ || !ctl_->ranks[r].attached.load()) continue;
        if (process_dead(ctl_->ranks[r].pid)) {
          dead_peer_.store(r);
          ALOG("[adapm] rank " << rank_ << ": peer rank " << r << " (pid " << ctl_->ranks[r].pid
                               << ") died - aborting all collective waits");
          ctl_->sync_barrier.broken.store(1);
          ctl_->node_barrier.broken.store(1);
          ctl_->worker_barrier.broken.store(1);
        }
      }
      for (int i = 0; i < period_ms / 10 && !fd_stop_.load(std::memory_order_acquire); ++i)
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
  });
}

void Fabric::stop_failure_detector() {
  fd_stop_.store(true, std::memory_order_release);
  if (fd_thread_.joinable()) fd_thread_.join();
}


}  // namespace adapm
