"""Reference-facing entry points (same module paths as wyhlovecpp/GPT-Image-Edit's `univa` package)
re-hosted over gpt_image_edit_b200 / libb2f.  Only the hot path's surface is provided."""
