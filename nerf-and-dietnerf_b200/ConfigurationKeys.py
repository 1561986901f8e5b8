"""YAML configuration keys read by the render/train hot path.

The key STRINGS are the config-file contract of the reference (src/ConfigurationKeys.py) and are kept verbatim so
the reference's config_files/*.yaml load unchanged; only the keys the hot path and its immediate callers read are
listed here.
"""
# top-level keys and blocks
EXISTING_SAVE_DIR_NAME = 'existing_save_dir_name'
STARTING_EPOCH_NUMBER = 'starting_epoch_number'
DATASET_TYPE = 'dataset_type'
DATASET_LOCATION = 'dataset_location'
PICS_INDICES_TO_USE_IN_DATASET = 'pics_indices_to_use_in_dataset'
GENERAL_SAVE_LOCATION = 'general_save_location'
TASKS_TO_PERFORM = 'tasks_to_perform'
START_TRAINING = 'start_training'
SAVE_DATASET_VIDEO = 'save_dataset_video'
RENDER_AND_SAVE_TEST_L_TO_R_VIDEO = 'render_and_save_test_left_to_right_video'
RENDER_AND_SAVE_TEST_SPHERE_VIDEO = 'render_and_save_test_sphere_video'
RENDER_AND_SAVE_TEST_PATH_VIDEO = 'render_and_save_test_path_video'
NEURAL_NET = 'neural_net'
RENDER = 'render'
TRAINING = 'training'
VIDEO = 'video'
BLENDER = 'blender'
COLMAP = 'colmap'

# neural_net:
TYPE_OF_MODEL = 'type_of_model'
HIDDEN_LAYER_DIM = 'hidden_layer_dim'
LAST_HIDDEN_LAYER_DIM = 'last_hidden_layer_dim'
LEAKY_RELU_ALPHA = 'leaky_relu_alpha'
N_POS_ENC_DIM_XYZ = 'n_pos_enc_dim_xyz'
N_POS_ENC_VIEW_DIR = 'n_pos_enc_view_dir'
N_ANGLES_FOR_MODEL = 'n_angles_for_model'
N_RAYS_IN_BATCH_TRAIN = 'n_rays_in_batch_train'
N_RAYS_IN_BATCH_RENDER = 'n_rays_in_batch_render'

# render:
N_RENDER_SAMPLES_COARSE = 'n_render_samples_coarse'
N_RENDER_SAMPLES_FINE = 'n_render_samples_fine'
NEAR_DEPTH_RENDER = 'near_depth_render'
FAR_DEPTH_RENDER = 'far_depth_render'

# training:
N_EPOCHS = 'n_epochs'
OPTIMIZER_LR = 'optimizer_lr'
TEST_IMG_IDX = 'test_img_idx'
IDX_TRAIN_IMG_TO_PLOT = 'idx_train_img_to_plot'

NERF_MODEL = 'NeRF'
DIETNERF_MODEL = 'DietNeRF'

# video:
FPS_TRAIN_SET_VIDEO = 'fps_train_set_video'
FPS_RENDER_VIDEO = 'fps_render_video'
IMG_INDICES_FOR_PATH_VIDEO = 'img_indices_for_path_video'
