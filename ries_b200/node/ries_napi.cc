// Node N-API addon: gives src/workers/ries.worker.js the object shape it already uses
// (createRIESModule({print, printErr}) -> {callMain}, /root/reference/src/workers/ries.worker.js:64-123,
// 226-234) on top of the C ABI of libries_b200.so (include/ries_b200.h).
//
// NOT built in this repository's image (node is not installed there): see INTEGRATION.md.  Build with
// node-gyp and the binding.gyp next to this file on a machine that has node, CUDA and a B200.
#include <napi.h>

#include <string>
#include <vector>

#include "ries_b200.h"

namespace {

void sink(void* ctx, const char* buf, size_t n) { static_cast<std::string*>(ctx)->append(buf, n); }

// callMain(args: string[]) -> { code: number, stdout: string }
Napi::Value CallMain(const Napi::CallbackInfo& info) {
  Napi::Env env = info.Env();
  if (info.Length() < 1 || !info[0].IsArray()) {
    Napi::TypeError::New(env, "callMain(args: string[])").ThrowAsJavaScriptException();
    return env.Null();
  }
  Napi::Array a = info[0].As<Napi::Array>();
  std::vector<std::string> args{"ries"};
  for (uint32_t i = 0; i < a.Length(); i++) args.push_back(a.Get(i).ToString().Utf8Value());
  std::vector<char*> argv;
  for (auto& s : args) argv.push_back(const_cast<char*>(s.c_str()));
  std::string out;
  int code = ries_main((int)argv.size(), argv.data(), sink, &out);
  Napi::Object r = Napi::Object::New(env);
  r.Set("code", Napi::Number::New(env, code));
  r.Set("stdout", Napi::String::New(env, out));
  return r;
}

Napi::Value EngineName(const Napi::CallbackInfo& info) { return Napi::String::New(info.Env(), ries_engine_name()); }

Napi::Object Init(Napi::Env env, Napi::Object exports) {
  exports.Set("callMain", Napi::Function::New(env, CallMain));
  exports.Set("engineName", Napi::Function::New(env, EngineName));
  return exports;
}

}  // namespace

NODE_API_MODULE(ries_b200, Init)
