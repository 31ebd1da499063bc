{ "targets": [ { "target_name": "zlibts_b200", "sources": ["addon.cc"],
    "libraries": ["-L<(module_root_dir)/../zlib.ts_b200", "-lzlibts_b200", "-Wl,-rpath,<(module_root_dir)/../zlib.ts_b200"] } ] }
