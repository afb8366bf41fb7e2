"""Drop-in for src/extractor/vf_extract.py - in the reference a byte-for-byte duplicate of src/video_frames_extract.py
(SURVEY.md section 1, L0); here one implementation, re-exported."""
from ..video_frames_extract import (extract_frames_general, extract_frames_residual, extract_frames_residual_yuv,   # noqa: F401
                                    extract_frames_yuv, process_video, process_video_residual, sample_clip, sample_video,
                                    sample_yuv420p, selected_indices)
