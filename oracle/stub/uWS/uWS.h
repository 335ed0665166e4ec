// TEST INFRASTRUCTURE (oracle). Transport stub, no arithmetic.
//
// The reference includes <uWS/uWS.h> (src/main.cpp:4) only for its websocket
// event loop (src/main.cpp:1157,1214-1216,1466,1471,1476-1494).  uWebSockets
// (pinned e94b6e1 by install-ubuntu.sh:5) is not in this image and carries no
// planner arithmetic, so this header provides just enough surface for the
// unmodified src/main.cpp to compile.  Hub::run() is defined by the harness
// (oracle/ref_harness.cpp) and drives the stored onMessage handler.
#pragma once
#include <cstddef>
#include <functional>
#include <string>

namespace uWS {

enum OpCode { TEXT = 1, BINARY = 2 };
enum { CLIENT = 0, SERVER = 1 };

struct HttpRequest {};

// Sink that receives whatever the handler passes to WebSocket::send.
struct SendSink {
  std::string last;
  long count = 0;
};
SendSink &send_sink();

template <int Role>
struct WebSocket {
  void send(const char *data, size_t length, OpCode) {
    SendSink &s = send_sink();
    s.last.assign(data, length);
    s.count++;
  }
  void close() {}
};

struct Hub {
  typedef std::function<void(WebSocket<SERVER>, char *, size_t, OpCode)> MessageFn;
  MessageFn message_fn;

  void onMessage(MessageFn fn) { message_fn = fn; }
  template <class F> void onConnection(F) {}
  template <class F> void onDisconnection(F) {}
  bool listen(const char *, int) { return true; }
  // Defined by the harness: feeds queued telemetry strings to message_fn and
  // then leaves by throwing (ref_main has no return after h.run()).
  void run();
};

}  // namespace uWS
