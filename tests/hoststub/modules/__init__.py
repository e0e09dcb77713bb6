"""Minimal stand-in for the AUTOMATIC1111 sdwui `modules` package — TEST INFRASTRUCTURE.

sdwui is not installable offline; this stub provides exactly the symbols the
reference extension (and our drop-in) import, so the plugin hooks can be driven end to end:
process_images() -> scripts.before_process -> process_images_inner -> postprocess_batch_list -> postprocess.
The master's own generation is delegated to `modules.processing.MASTER_GENERATOR` (tests plug an oracle or
a constant-image function in).
"""
