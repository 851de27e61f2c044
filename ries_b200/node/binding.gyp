{
  "targets": [{
    "target_name": "ries_b200",
    "sources": ["ries_napi.cc"],
    "include_dirs": ["<!@(node -p \"require('node-addon-api').include\")", "../../include"],
    "defines": ["NAPI_DISABLE_CPP_EXCEPTIONS"],
    "libraries": ["-L<(module_root_dir)/..", "-lries_b200", "-Wl,-rpath,<(module_root_dir)/.."]
  }]
}
