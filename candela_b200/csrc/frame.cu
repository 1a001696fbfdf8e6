// frame-level calls (filled in below)
