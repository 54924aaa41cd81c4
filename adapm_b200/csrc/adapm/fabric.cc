#include "fabric.h"

#include <dirent.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <poll.h>
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

// ---- passing file descriptors between the ranks of a box (VMM heap handles, the multicast object): one unix
// datagram socket per rank in the abstract namespace (nothing to clean up on the file system), SCM_RIGHTS messages
struct FdMsg { int32_t sender; int32_t kind; };
enum FdKind : int32_t { FD_HEAP = 1, FD_MULTICAST = 2 };

sockaddr_un fdx_addr(const std::string& name, socklen_t* len) {
  sockaddr_un a;
  memset(&a, 0, sizeof(a));
  a.sun_family = AF_UNIX;
  const size_t n = std::min(name.size(), sizeof(a.sun_path) - 2);
  memcpy(a.sun_path + 1, name.data(), n);   // sun_path[0] == 0: abstract socket
  *len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + n);
  return a;
}
int fdx_bind(const std::string& name) {
  int s = socket(AF_UNIX, SOCK_DGRAM | SOCK_CLOEXEC, 0);
  ADAPM_CHECK(s >= 0, "socket(AF_UNIX) failed: " << strerror(errno));
  socklen_t len;
  sockaddr_un a = fdx_addr(name, &len);
  ADAPM_CHECK(bind(s, (sockaddr*)&a, len) == 0, "bind(" << name << ") failed: " << strerror(errno));
  return s;
}
void fdx_send(int sock, const std::string& to, int fd, FdMsg msg) {
  socklen_t len;
  sockaddr_un a = fdx_addr(to, &len);
  iovec iov{&msg, sizeof(msg)};
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  memset(ctl, 0, sizeof(ctl));
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_name = &a; mh.msg_namelen = len;
  mh.msg_iov = &iov; mh.msg_iovlen = 1;
  mh.msg_control = ctl; mh.msg_controllen = sizeof(ctl);
  cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  for (int attempt = 0;; ++attempt) {
    if (sendmsg(sock, &mh, 0) == (ssize_t)sizeof(msg)) return;
    ADAPM_CHECK((errno == EAGAIN || errno == ENOBUFS || errno == ECONNREFUSED || errno == ENOENT) && attempt < 2000,
                "sendmsg(fd) to " << to << " failed: " << strerror(errno));
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
}
int fdx_recv(int sock, FdMsg* msg, double timeout_s) {
  pollfd pf{sock, POLLIN, 0};
  int pr = poll(&pf, 1, (int)(timeout_s * 1000));
  ADAPM_CHECK(pr > 0, "timed out waiting for a peer's memory handle");
  iovec iov{msg, sizeof(*msg)};
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  msghdr mh;
  memset(&mh, 0, sizeof(mh));
  mh.msg_iov = &iov; mh.msg_iovlen = 1;
  mh.msg_control = ctl; mh.msg_controllen = sizeof(ctl);
  ssize_t n = recvmsg(sock, &mh, MSG_CMSG_CLOEXEC);
  ADAPM_CHECK(n == (ssize_t)sizeof(*msg), "recvmsg(fd) failed: " << strerror(errno));
  cmsghdr* cm = CMSG_FIRSTHDR(&mh);
  ADAPM_CHECK(cm && cm->cmsg_level == SOL_SOCKET && cm->cmsg_type == SCM_RIGHTS, "peer message carries no file descriptor");
  int fd = -1;
  memcpy(&fd, CMSG_DATA(cm), sizeof(int));
  return fd;
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
    if (cuda_ && !vmm_.empty()) {
      cudamem::set_device(device_);
      if (mc_handle_) cudamem::mc_unmap(mc_handle_, device_, mc_heap_, vmm_[rank_].size);
      for (auto& h : vmm_) cudamem::vmm_free(h);
      for (int r = 0; r < world_; ++r) heaps_[r] = nullptr;
    }
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
    if (cuda_ && allocate_vmm_heaps(b)) return;
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

  // Heaps as VMM allocations that peers import through POSIX file descriptors, bound to one NVLS multicast object
  // when every device of the job supports it (NVSwitch boxes): mc_heap() + off then addresses offset `off` of ALL
  // heaps at once. Returns false (nothing allocated) when some rank lacks VMM support or ADAPM_VMM=0: the caller
  // falls back to cudaMalloc + CUDA IPC.
  bool allocate_vmm_heaps(uint64_t b) {
    RankControl& me = ctl_->ranks[rank_];
    cudamem::set_device(device_);
    bool mc = false;
    bool vmm = cudamem::vmm_supported(device_, &mc);
    if (const char* e = getenv("ADAPM_VMM")) vmm = vmm && atoi(e) != 0;
    if (const char* e = getenv("ADAPM_MULTICAST")) mc = mc && atoi(e) != 0;
    me.vmm_ok = vmm ? 1 : 0;
    me.mc_ok = (vmm && mc) ? 1 : 0;
    me.device = device_;
    node_barrier("vmm capabilities");
    bool all_vmm = true, all_mc = world_ > 1;
    for (int r = 0; r < world_; ++r) { all_vmm = all_vmm && ctl_->ranks[r].vmm_ok; all_mc = all_mc && ctl_->ranks[r].mc_ok; }
    for (int r = 0; r < world_ && all_mc; ++r)     // one device per rank: a device cannot join a multicast object twice
      for (int q = 0; q < r; ++q) if (ctl_->ranks[q].device == ctl_->ranks[r].device) all_mc = false;
    if (!all_vmm) return false;
    const uint64_t gran = cudamem::vmm_granularity(device_, world_, all_mc);
    const uint64_t size = (b + gran - 1) / gran * gran;
    vmm_.assign(world_, cudamem::VmmHeap());
    vmm_[rank_] = cudamem::vmm_alloc(device_, size);
    heaps_[rank_] = vmm_[rank_].va;
    me.heap_bytes = size;
    heap_bytes_[rank_] = size;
    auto sock_name = [&](int r) { return prefix_.substr(1) + "_fd" + std::to_string(r); };
    const int sock = fdx_bind(sock_name(rank_));
    node_barrier("vmm sockets bound");
    for (int r = 0; r < world_; ++r) {
      if (r == rank_) continue;
      const int fd = cudamem::vmm_export_fd(vmm_[rank_]);
      fdx_send(sock, sock_name(r), fd, FdMsg{rank_, FD_HEAP});
      close(fd);
    }
    if (all_mc && rank_ == 0) {
      int fd = -1;
      mc_handle_ = cudamem::mc_create(world_, size, &fd);
      for (int r = 1; r < world_; ++r) fdx_send(sock, sock_name(r), fd, FdMsg{0, FD_MULTICAST});
      close(fd);
    }
    int expect = (world_ - 1) + ((all_mc && rank_ != 0) ? 1 : 0);
    while (expect-- > 0) {
      FdMsg m;
      const int fd = fdx_recv(sock, &m, timeout_s_);
      if (m.kind == FD_MULTICAST) {
        mc_handle_ = cudamem::mc_import_fd(fd);
      } else {
        ADAPM_CHECK(m.kind == FD_HEAP && m.sender >= 0 && m.sender < world_ && m.sender != rank_, "unexpected handle message");
        vmm_[m.sender] = cudamem::vmm_import_fd(device_, fd, size);
        heaps_[m.sender] = vmm_[m.sender].va;
        heap_bytes_[m.sender] = size;
      }
    }
    close(sock);
    if (all_mc) {
      cudamem::mc_add_device(mc_handle_, device_);
      node_barrier("multicast devices added");
      mc_heap_ = cudamem::mc_bind_and_map(mc_handle_, device_, vmm_[rank_]);
    }
    node_barrier("vmm heaps mapped");
    VLOG(1, "rank " << rank_ << ": VMM heap " << (size >> 20) << " MiB" << (mc_heap_ ? ", bound to an NVLS multicast object" : ""));
    return true;
  }

 private:
  std::string prefix_;
  size_t ctl_sz_ = 0;
  std::vector<uint64_t> heap_bytes_;
  bool my_heap_created_ = false;
  std::vector<cudamem::VmmHeap> vmm_;
  unsigned long long mc_handle_ = 0;
};

}  // namespace

// Self-test of the file-descriptor channel (no GPU needed): a pipe's read end travels from one abstract unix socket to
// another; the bytes written into the pipe must come out of the received descriptor.
bool fabric_fdpass_selftest() {
  const std::string a = "adapm_selftest_a_" + std::to_string((long)getpid()), b = "adapm_selftest_b_" + std::to_string((long)getpid());
  int sa = fdx_bind(a), sb = fdx_bind(b);
  int p[2];
  if (pipe(p) != 0) return false;
  fdx_send(sa, b, p[0], FdMsg{7, FD_HEAP});
  FdMsg m;
  int fd = fdx_recv(sb, &m, 5.0);
  const char msg[] = "adapm";
  bool ok = write(p[1], msg, sizeof(msg)) == (ssize_t)sizeof(msg);
  char buf[16] = {0};
  ok = ok && read(fd, buf, sizeof(msg)) == (ssize_t)sizeof(msg) && memcmp(buf, msg, sizeof(msg)) == 0;
  ok = ok && m.sender == 7 && m.kind == FD_HEAP && fd != p[0];
  close(fd); close(p[0]); close(p[1]); close(sa); close(sb);
  return ok;
}

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
