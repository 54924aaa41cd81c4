// Host-side RPC between ranks: an (int head, byte-string body) request/response channel.
//
// Capability parity with the legacy PS-Lite `SimpleApp` (include/ps/simple_app.h:32-184: Request / Response /
// Wait, request and response handles) and the request tracking of `Customer`
// (include/ps/internal/customer.h:24-110). The parameter manager itself never uses it (its traffic is loads and
// reductions on the symmetric heap); it exists for application-level control messages and for the legacy
// KVWorker/KVServer API (adapm_b200/legacy.py).
//
// Design: every rank owns one multi-producer / single-consumer ring of fixed-size slots inside the shared
// control block. A sender claims a ticket with fetch_add, waits until the slot is free, writes the fragment and
// publishes it with a release store; the owner's router thread drains the ring in ticket order, reassembles
// fragmented bodies and dispatches to the endpoint registered under (app_id, customer_id).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "base.h"
#include "control.h"

namespace adapm {

// recv_id of Request(): a rank, or one of the groups (every rank is a server and hosts workers)
constexpr int kServerGroup = -2;
constexpr int kWorkerGroup = -4;
constexpr int kAllNodes = -6;

struct SimpleData {
  int head = 0;
  std::string body;
  int sender = 0;        // rank of the sender
  int timestamp = 0;
  int customer_id = 0;   // sender's customer id
};

class Server;
class SimpleApp;

// One per Server (created on first use): drains this rank's mailbox and dispatches.
class MailRouter {
 public:
  explicit MailRouter(Server* server);
  ~MailRouter();
  void stop();
  void attach(SimpleApp* app);
  void detach(SimpleApp* app);
  // Fragment + deliver one message into rank `to`'s mailbox (or the local queue when to == my rank).
  void send(int to, int app_id, int customer_id, int head, int timestamp, bool request, const std::string& body);

 private:
  struct Msg { int app_id, customer_id; bool request; SimpleData d; };
  void loop();
  void dispatch(Msg&& m);
  // defer = true: completed messages are queued (local_) instead of dispatched - used while the router thread itself
  // waits for a free slot in a peer's mailbox, so that two ranks answering each other never block each other
  bool drain_ring(bool defer = false);

  Server* server_;
  std::thread thread_;
  std::atomic<bool> stop_{false};
  std::mutex mu_;                         // endpoints_, pending_, local_
  std::map<std::pair<int, int>, SimpleApp*> endpoints_;
  std::deque<Msg> pending_;               // messages for endpoints that are not registered yet
  std::deque<Msg> local_;                 // self-sends
  std::map<std::pair<int, uint32_t>, Msg> partial_;   // (sender, msg_id) -> message being reassembled
  std::atomic<uint32_t> next_msg_id_{1};
};

class SimpleApp {
 public:
  using Handle = std::function<void(const SimpleData& recved, SimpleApp* app)>;

  // serves_requests = false: a pure client endpoint (it only receives the responses to its own requests)
  SimpleApp(int app_id, int customer_id, Server& server, bool serves_requests = true);
  virtual ~SimpleApp();

  // Sends (head, body) to `recv_id` (rank or group); returns the timestamp to Wait() on.
  int Request(int req_head, const std::string& req_body, int recv_id);
  // Blocks until every receiver of request `timestamp` has responded (watchdog: wait_timeout_s).
  void Wait(int timestamp);
  int NumResponse(int timestamp);
  // Answers a received request.
  void Response(const SimpleData& recv_req, const std::string& res_body = "");

  void set_request_handle(const Handle& h) { std::lock_guard<std::mutex> lk(mu_); request_handle_ = h; }
  void set_response_handle(const Handle& h) { std::lock_guard<std::mutex> lk(mu_); response_handle_ = h; }
  int app_id() const { return app_id_; }
  int customer_id() const { return customer_id_; }
  bool serves_requests() const { return serves_requests_; }
  Server& server() { return *server_; }

 private:
  friend class MailRouter;
  void on_message(bool request, const SimpleData& d);   // router thread

  Server* server_;
  int app_id_, customer_id_;
  bool serves_requests_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::vector<std::pair<int, int>> tracker_;   // per timestamp: (expected, received)
  Handle request_handle_, response_handle_;
};

}  // namespace adapm
