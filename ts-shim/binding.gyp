{
  "targets": [
    {
      "target_name": "swfr_b200",
      "sources": ["native/addon.cc"],
      "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "<(module_root_dir)/../include"],
      "defines": ["NAPI_CPP_EXCEPTIONS"],
      "cflags_cc": ["-std=c++17", "-fexceptions"],
      "libraries": ["-L<(module_root_dir)/../swf_renderer_b200", "-lswfr_b200", "-Wl,-rpath,<(module_root_dir)/../swf_renderer_b200"]
    }
  ]
}
