"""qwen3_rs_b200 -- B200-native (sm_100a) drop-in for qwen3-rs's quantized forward pass."""
