// Declarations-only stand-in for node-addon-api's <napi.h>: just the surface ts-shim/src/addon.cc uses, so that the
// addon can be type-checked against include/swfr.h (g++ -fsyntax-only) in an image without Node.  Test infrastructure.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <initializer_list>

enum napi_typedarray_type { napi_uint8_array, napi_uint8_clamped_array, napi_int32_array };

namespace Napi {
class Env;
class Value;
class Object;
class Boolean;
class Number;
class String;
class Array;
class Function;

class Env {
 public:
  Value Undefined() const;
};

class Value {
 public:
  bool IsUndefined() const;
  bool IsNull() const;
  Boolean ToBoolean() const;
  template <class T>
  T As() const;
  Env Env_() const;
};

class Boolean : public Value {
 public:
  bool Value() const;
};

class Number : public Value {
 public:
  static Number New(Env env, double v);
  int32_t Int32Value() const;
  uint32_t Uint32Value() const;
};

class String : public Value {
 public:
  std::string Utf8Value() const;
};

class Object : public Value {
 public:
  bool Has(const char *key) const;
  Value Get(const char *key) const;
  Value Get(uint32_t index) const;
  template <class V>
  void Set(const char *key, const V &v);
};

class Array : public Object {
 public:
  uint32_t Length() const;
};

class ArrayBuffer : public Object {
 public:
  static ArrayBuffer New(Env env, size_t bytes);
  void *Data();
};

template <class T>
class TypedArrayOf : public Object {
 public:
  static TypedArrayOf New(Env env, size_t n);
  static TypedArrayOf New(Env env, size_t n, ArrayBuffer buf, size_t offset, napi_typedarray_type type);
  T *Data() const;
  size_t ElementLength() const;
  size_t ByteLength() const;
  T operator[](size_t i) const;
};
using Uint8Array = TypedArrayOf<uint8_t>;
using Int32Array = TypedArrayOf<int32_t>;

class Error {
 public:
  static Error New(Env env, const std::string &msg);
  static Error New(Env env, const char *msg);
};

class CallbackInfo {
 public:
  Napi::Env Env() const;
  size_t Length() const;
  Value operator[](size_t i) const;
};

class Function : public Object {};

template <class T>
class ObjectWrap {
 public:
  explicit ObjectWrap(const CallbackInfo &info);
  virtual ~ObjectWrap();
  struct PropertyDescriptor {};
  using Method = Value (T::*)(const CallbackInfo &);
  static PropertyDescriptor InstanceMethod(const char *name, Method m);
  static Function DefineClass(Env env, const char *name, std::initializer_list<PropertyDescriptor> props);
};
}  // namespace Napi

#define NODE_API_MODULE(name, init) \
  Napi::Object napi_module_register_##name(Napi::Env env, Napi::Object exports) { return init(env, exports); }
