#include "rpc.h"

#include <chrono>
#include <cstring>

#include "log.h"
#include "node.h"

namespace adapm {

// ------------------------------------------------------------------------------------- router
MailRouter::MailRouter(Server* server) : server_(server) {
  thread_ = std::thread([this] {
    try {
      loop();
    } catch (const std::exception& e) {
      ALOG("[adapm] rank " << server_->my_rank() << " rpc router died: " << e.what());
    }
  });
}

MailRouter::~MailRouter() { stop(); }

void MailRouter::stop() {
  if (stop_.exchange(true)) return;
  if (thread_.joinable()) thread_.join();
}

void MailRouter::attach(SimpleApp* app) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto key = std::make_pair(app->app_id(), app->customer_id());
    ADAPM_CHECK(!endpoints_.count(key), "SimpleApp: (app_id, customer_id) already registered on this rank");
    endpoints_[key] = app;
    // messages that arrived before the endpoint existed go through the local queue again
    for (auto it = pending_.begin(); it != pending_.end();) {
      if (it->app_id == key.first) {
        local_.push_back(std::move(*it));
        it = pending_.erase(it);
      } else {
        ++it;
      }
    }
  }
}

void MailRouter::detach(SimpleApp* app) {
  std::lock_guard<std::mutex> lk(mu_);
  endpoints_.erase(std::make_pair(app->app_id(), app->customer_id()));
}

void MailRouter::send(int to, int app_id, int customer_id, int head, int timestamp, bool request,
                      const std::string& body) {
  const int me = server_->my_rank();
  if (to == me) {
    Msg m;
    m.app_id = app_id; m.customer_id = customer_id; m.request = request;
    m.d.head = head; m.d.body = body; m.d.sender = me; m.d.timestamp = timestamp; m.d.customer_id = customer_id;
    std::lock_guard<std::mutex> lk(mu_);
    local_.push_back(std::move(m));
    return;
  }
  ADAPM_CHECK(to >= 0 && to < server_->num_servers(), "rpc: receiver rank out of range");
  Mailbox& mb = server_->control()->mail[to];
  const uint32_t msg_id = next_msg_id_.fetch_add(1);
  const uint32_t total = (uint32_t)body.size();
  const double timeout = server_->options().wait_timeout_s;
  uint32_t off = 0;
  do {
    const uint32_t n = std::min<uint32_t>(MAIL_BODY, total - off);
    const uint64_t ticket = mb.tail.fetch_add(1, std::memory_order_acq_rel);
    MailSlot& s = mb.slots[ticket % MAIL_SLOTS];
    auto t0 = std::chrono::steady_clock::now();
    int spins = 0;
    // my turn on this slot: the consumer has passed ticket - MAIL_SLOTS and released the slot
    const bool on_router = std::this_thread::get_id() == thread_.get_id();
    while (mb.head.load(std::memory_order_acquire) + MAIL_SLOTS <= ticket || s.state.load(std::memory_order_acquire) != 0) {
      // a handler (router thread) that waits for a slot keeps emptying its own mailbox: the peer may be waiting for one
      if (on_router) drain_ring(/*defer=*/true);
      if (++spins < 100) std::this_thread::yield();
      else std::this_thread::sleep_for(std::chrono::microseconds(50));
      if ((spins & 1023) == 0) {
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        ADAPM_CHECK(el < timeout, "watchdog: rpc send timed out (receiver's mailbox stays full: peer dead?)");
      }
    }
    s.sender = me; s.app_id = app_id; s.customer_id = customer_id; s.head = head; s.timestamp = timestamp;
    s.flags = request ? 1 : 0;
    s.msg_id = msg_id; s.frag_off = off; s.total_len = total; s.frag_len = n;
    if (n) memcpy(s.body, body.data() + off, n);
    s.state.store(1, std::memory_order_release);
    off += n;
  } while (off < total);
}

bool MailRouter::drain_ring(bool defer) {
  Mailbox& mb = server_->control()->mail[server_->my_rank()];
  bool any = false;
  for (;;) {
    const uint64_t h = mb.head.load(std::memory_order_relaxed);
    MailSlot& s = mb.slots[h % MAIL_SLOTS];
    if (s.state.load(std::memory_order_acquire) != 1) break;
    any = true;
    auto key = std::make_pair((int)s.sender, (uint32_t)s.msg_id);
    Msg* m;
    auto it = partial_.find(key);
    if (it == partial_.end()) {
      Msg fresh;
      fresh.app_id = s.app_id; fresh.customer_id = s.customer_id; fresh.request = (s.flags & 1) != 0;
      fresh.d.head = s.head; fresh.d.sender = s.sender; fresh.d.timestamp = s.timestamp;
      fresh.d.customer_id = s.customer_id;
      fresh.d.body.reserve(s.total_len);
      m = &partial_.emplace(key, std::move(fresh)).first->second;
    } else {
      m = &it->second;
    }
    m->d.body.append(s.body, s.frag_len);
    const bool complete = m->d.body.size() >= s.total_len;
    s.state.store(0, std::memory_order_release);
    mb.head.store(h + 1, std::memory_order_release);
    if (complete) {
      Msg done = std::move(*m);
      partial_.erase(key);
      if (defer) {
        std::lock_guard<std::mutex> lk(mu_);
        local_.push_back(std::move(done));
      } else {
        dispatch(std::move(done));
      }
    }
  }
  return any;
}

void MailRouter::dispatch(Msg&& m) {
  SimpleApp* app = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu_);
    // requests go to the serving endpoint of that app on this rank (lowest customer id); responses go back to
    // exactly the requesting customer
    if (m.request) {
      for (auto it = endpoints_.lower_bound(std::make_pair(m.app_id, INT32_MIN));
           it != endpoints_.end() && it->first.first == m.app_id; ++it) {
        if (it->second->serves_requests()) { app = it->second; break; }
      }
    } else {
      auto it = endpoints_.find(std::make_pair(m.app_id, m.customer_id));
      if (it != endpoints_.end()) app = it->second;
    }
    if (!app) {
      pending_.push_back(std::move(m));
      return;
    }
  }
  app->on_message(m.request, m.d);
}

void MailRouter::loop() {
  int idle = 0;
  while (!stop_.load(std::memory_order_acquire)) {
    bool any = drain_ring();
    for (;;) {
      Msg m;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (local_.empty()) break;
        m = std::move(local_.front());
        local_.pop_front();
      }
      any = true;
      dispatch(std::move(m));
    }
    if (any) { idle = 0; continue; }
    if (++idle < 50) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(idle < 2000 ? 50 : 500));
  }
}

// ------------------------------------------------------------------------------------- endpoint
SimpleApp::SimpleApp(int app_id, int customer_id, Server& server, bool serves_requests)
    : server_(&server), app_id_(app_id), customer_id_(customer_id), serves_requests_(serves_requests) {
  // defaults of the reference: a request is acknowledged with an empty response, responses are only counted
  // (simple_app.h:100-110)
  request_handle_ = [](const SimpleData& req, SimpleApp* app) { app->Response(req); };
  response_handle_ = [](const SimpleData&, SimpleApp*) {};
  server_->router().attach(this);
}

SimpleApp::~SimpleApp() {
  if (MailRouter* r = server_->router_if_any()) r->detach(this);
}

int SimpleApp::Request(int req_head, const std::string& req_body, int recv_id) {
  const int world = server_->num_servers();
  std::vector<int> to;
  if (recv_id >= 0) {
    ADAPM_CHECK(recv_id < world, "SimpleApp::Request: receiver rank out of range");
    to.push_back(recv_id);
  } else {
    ADAPM_CHECK(recv_id == kServerGroup || recv_id == kWorkerGroup || recv_id == kAllNodes,
                "SimpleApp::Request: unknown receiver group");
    for (int r = 0; r < world; ++r) to.push_back(r);
  }
  int ts;
  {
    std::lock_guard<std::mutex> lk(mu_);
    ts = (int)tracker_.size();
    tracker_.emplace_back((int)to.size(), 0);
  }
  for (int r : to) server_->router().send(r, app_id_, customer_id_, req_head, ts, true, req_body);
  return ts;
}

void SimpleApp::Response(const SimpleData& req, const std::string& res_body) {
  server_->router().send(req.sender, app_id_, req.customer_id, req.head, req.timestamp, false, res_body);
}

int SimpleApp::NumResponse(int timestamp) {
  std::lock_guard<std::mutex> lk(mu_);
  ADAPM_CHECK(timestamp >= 0 && (size_t)timestamp < tracker_.size(), "SimpleApp: unknown timestamp");
  return tracker_[timestamp].second;
}

void SimpleApp::Wait(int timestamp) {
  std::unique_lock<std::mutex> lk(mu_);
  ADAPM_CHECK(timestamp >= 0 && (size_t)timestamp < tracker_.size(), "SimpleApp::Wait: unknown timestamp");
  const auto deadline = std::chrono::steady_clock::now() +
                        std::chrono::duration_cast<std::chrono::steady_clock::duration>(
                            std::chrono::duration<double>(server_->options().wait_timeout_s));
  while (tracker_[timestamp].second < tracker_[timestamp].first) {
    if (cv_.wait_until(lk, deadline) == std::cv_status::timeout)
      ADAPM_CHECK(tracker_[timestamp].second >= tracker_[timestamp].first,
                  "watchdog: SimpleApp::Wait timed out (a receiver never responded)");
  }
}

void SimpleApp::on_message(bool request, const SimpleData& d) {
  Handle h;
  {
    std::lock_guard<std::mutex> lk(mu_);
    h = request ? request_handle_ : response_handle_;
  }
  if (h) h(d, this);
  if (!request) {
    std::lock_guard<std::mutex> lk(mu_);
    if (d.timestamp >= 0 && (size_t)d.timestamp < tracker_.size()) ++tracker_[d.timestamp].second;
    cv_.notify_all();
  }
}

}  // namespace adapm
