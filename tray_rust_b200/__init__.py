"""tray_rust_b200 — B200-native render hot path for tray_rust."""
