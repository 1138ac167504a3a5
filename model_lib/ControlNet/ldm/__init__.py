"""Same dotted path as the reference (Boese0601/MagicDance) — see INTEGRATION.md."""
