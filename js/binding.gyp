{
  "targets": [{
    "target_name": "manatee_gpu",
    "sources": ["src/binding.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../manatee_b200", "-lmanatee_gpu",
                  "-Wl,-rpath,<(module_root_dir)/../manatee_b200"]
  }]
}
